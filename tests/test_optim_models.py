import copy

import pytest
import torch

from atomo_b200.models import build_model, NETWORKS, LeNetSplit, FC_NN_Split, ResNetSplit18, input_shape
from atomo_b200.optim import SGD, Adam


@pytest.mark.parametrize("kw", [dict(momentum=0.0), dict(momentum=0.9), dict(momentum=0.9, nesterov=True),
                                dict(momentum=0.5, weight_decay=1e-2), dict(momentum=0.9, dampening=0.1)])
def test_sgd_external_grads_matches_torch(kw):
    torch.manual_seed(0)
    a = torch.nn.Linear(7, 5)
    b = copy.deepcopy(a)
    ours, ref = SGD(a.parameters(), lr=0.1, **kw), torch.optim.SGD(b.parameters(), lr=0.1, **kw)
    for _ in range(4):
        grads = [torch.randn_like(p) for p in a.parameters()]
        for p, g in zip(b.parameters(), grads):
            p.grad = g.clone()
        ours.step(grads=[g.numpy() for g in grads])  # numpy accepted like optim/sgd.py:74
        ref.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-6)
    ours.set_lr(0.01)
    assert ours.param_groups[0]["lr"] == 0.01


@pytest.mark.parametrize("amsgrad", [False, True])
def test_adam_external_grads_matches_torch(amsgrad):
    torch.manual_seed(0)
    a = torch.nn.Linear(6, 3)
    b = copy.deepcopy(a)
    ours = Adam(a.parameters(), lr=1e-2, amsgrad=amsgrad, weight_decay=1e-3)
    ref = torch.optim.Adam(b.parameters(), lr=1e-2, amsgrad=amsgrad, weight_decay=1e-3)
    for _ in range(5):
        grads = [torch.randn_like(p) for p in a.parameters()]
        for p, g in zip(b.parameters(), grads):
            p.grad = g.clone()
        ours.step(grads=grads)
        ref.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-5)


@pytest.mark.parametrize("net,params,tensors", [("LeNet", 431080, 8), ("FC", 1033510, 6),
                                                ("ResNet18", 11173962, 62), ("VGG11", 9756426, 38)])
def test_param_counts_match_reference_models(net, params, tensors):
    m = build_model(net)
    ps = list(m.parameters())
    assert sum(p.numel() for p in ps) == params and len(ps) == tensors


@pytest.mark.parametrize("net", ["LeNet", "FC", "ResNet18", "ResNet34", "ResNet50", "VGG11", "VGG16", "DenseNetSmall"])
def test_forward_shapes(net):
    m = build_model(net, 10).eval()
    x = torch.randn(2, *input_shape(net))
    assert m(x).shape == (2, 10)


def test_resnet50_imagenet_stem_and_alexnet():
    m = build_model("ResNet50", 1000, "ImageNet").eval()
    assert m(torch.randn(1, 3, 64, 64)).shape == (1, 1000)
    a = build_model("AlexNet", 10).eval()
    assert a(torch.randn(1, 3, 227, 227)).shape == (1, 10)
    with pytest.raises(ValueError):
        build_model("NoSuchNet")
    assert "DenseNet" in NETWORKS


def test_split_models_emit_every_gradient_and_match_autograd():
    torch.manual_seed(0)
    for ctor, shape in ((LeNetSplit, (4, 1, 28, 28)), (FC_NN_Split, (4, 1, 28, 28)), (ResNetSplit18, (2, 3, 32, 32))):
        sm = ctor()
        ref = copy.deepcopy(sm)
        sm.train(); ref.train()
        x, y = torch.randn(*shape), torch.randint(0, 10, (shape[0],))
        loss = sm.criterion(sm(x), y)
        got = {}
        sm.backward(loss, emit=lambda i, p, g: got.__setitem__(i, g.clone()))
        assert sorted(got) == list(range(len(list(sm.parameters()))))
        # same gradients as a plain end-to-end backward
        out = x
        for st in ref.stages:
            out = st(out)
        ref.criterion(out, y).backward()
        for (i, g), p in zip(sorted(got.items()), ref.parameters()):
            assert torch.allclose(g, p.grad, atol=1e-5), (ctor.__name__, i)


def test_split_kill_variants():
    sm = LeNetSplit()
    x, y = torch.randn(2, 1, 28, 28), torch.randint(0, 10, (2,))
    calls = []
    loss = sm.criterion(sm(x), y)
    killed = sm.backward_signal_kill(loss, emit=lambda i, p, g: calls.append(i), kill_signal=lambda: len(calls) >= 2)
    assert killed and 0 < len(calls) < 8
    loss = sm.criterion(sm(x), y)
    assert sm.backward_timeout_kill(loss, timeout_s=-1.0) is True
    loss = sm.criterion(sm(x), y)
    sm.backward_single(loss)


def test_sidecar_restores_the_full_adam_state(tmp_path):
    """--resume on the role path: the `_optim` sidecar carries the optimizer's whole state (Adam moments, step
    counts), so a resumed PS takes exactly the step the uninterrupted one would have taken."""
    from atomo_b200.optim import Adam
    from atomo_b200.utils import checkpoint as ckpt
    torch.manual_seed(0)
    d = str(tmp_path) + "/"

    def make():
        torch.manual_seed(1)
        net = torch.nn.Linear(6, 3)
        return net, Adam(net.parameters(), lr=0.01, amsgrad=True)

    grads = [[torch.randn(3, 6), torch.randn(3)] for _ in range(5)]
    a, oa = make()
    for g in grads[:3]:
        oa.step(grads=g)
    ckpt.save_model(d, 3, a)
    ckpt.save_sidecar(d, 3, oa, lr=0.01)
    for g in grads[3:]:
        oa.step(grads=g)
    b, ob = make()
    ckpt.load_model(d, 3, b)
    side = ckpt.load_sidecar(d, 3, ob)
    assert side["step"] == 3 and side["lr"] == 0.01
    for g in grads[3:]:
        ob.step(grads=g)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-7)
