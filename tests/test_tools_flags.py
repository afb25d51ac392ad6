"""Ops tooling + flag surface."""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_flag_surface_and_defaults():
    from atomo_b200.utils.flags import add_fit_args, bool_flag
    a = add_fit_args(argparse.ArgumentParser(), [])
    ref_defaults = {  # distributed_nn.py:37-80
        "batch_size": 128, "test_batch_size": 1000, "max_steps": 10000, "epochs": 100, "lr": 0.01, "momentum": 0.5,
        "lr_shrinkage": 0.95, "no_cuda": False, "seed": 1, "log_interval": 10, "network": "LeNet", "code": "sgd",
        "bucket_size": 512, "dataset": "MNIST", "comm_type": "Bcast", "eval_freq": 50,
        "train_dir": "output/models/", "compress": False, "enable_gpu": False, "svd_rank": 0,
        "quantization_level": 4,
    }
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, k
    assert a.num_aggregate == 0  # documented divergence: the flag now has an effect, 0 = all workers
    # the reference's `type=bool` contract: scripts pass `--enable-gpu=` for False, anything else is True
    b = add_fit_args(argparse.ArgumentParser(), ["--enable-gpu=", "--compress=yes"])
    assert b.enable_gpu is False and b.compress is True
    assert bool_flag("0") is False and bool_flag("False") is False and bool_flag("1") is True


def test_cluster_tool_command_map_and_cfg(tmp_path, monkeypatch):
    cl = _load(os.path.join(ROOT, "tools", "cluster.py"), "cluster_tool")
    # the reference's command names (pytorch_ec2.py:938-951)
    for cmd in ("launch", "get_hosts", "shutdown", "kill_all_python", "kill_python", "run_command", "setup_nfs",
                "list_idle_instances", "list_running_instances", "clean_launch_and_run"):
        assert cmd in cl.command_map
    cfg = cl.Cfg({"a": "x", "b": "%(a)s/y", "c": "%(b)s/z"})
    assert cfg["c"] == "x/y/z"
    rc, out = cl._run("localhost", "echo hello")
    assert rc == 0 and "hello" in out
    monkeypatch.setattr(cl.os.path, "dirname", lambda p: str(tmp_path))
    args = argparse.Namespace(hosts="n1,n2,n3", hostfile="")
    assert cl.get_hosts(args) == ["n1", "n2", "n3"]
    assert open(os.path.join(str(tmp_path), "hosts_address")).read().split() == ["n1", "n2", "n3"]
    assert "deeplearning-worker2" in open(os.path.join(str(tmp_path), "hosts_alias")).read()


def test_data_prepare_and_sv_decay_tools(tmp_path, capsys):
    from atomo_b200.data.data_prepare import main as prep
    prep(["--root", str(tmp_path), "--materialize", "8"])
    assert os.path.exists(os.path.join(str(tmp_path), "mnist_synthetic.pt"))
    sv = _load(os.path.join(ROOT, "tools", "sv_decay.py"), "sv_decay_tool")
    rows = sv.main(["--network", "LeNet", "--layer", "conv2.weight", "--steps", "2", "--every", "1",
                    "--batch-size", "8"])
    assert len(rows) == 3 and len(rows[0][1]) == 50 and rows[0][1][0] == 1.0  # (500, 50) matricization


def test_p2p_launcher_engine_selection(monkeypatch):
    """--backend p2p: bf16 + svd|qsvd|sgd runs the overlapped sharded engine, everything else the fp32-flat engine;
    asking the fp32-flat engine for Adam is refused instead of silently training with SGD."""
    import argparse
    import sys
    import types
    from atomo_b200.runtime import p2p_launcher as L
    from atomo_b200.utils.flags import add_fit_args

    made = []

    class Fake:
        def __init__(self, *a, **kw):
            made.append((type(self).__name__, kw))

    shadow = types.ModuleType("atomo_b200.runtime.shadow_engine")
    shadow.ShadowEngine = type("ShadowEngine", (Fake,), {})
    fused = types.ModuleType("atomo_b200.runtime.engine")
    fused.FusedEngine = type("FusedEngine", (Fake,), {})
    monkeypatch.setitem(sys.modules, "atomo_b200.runtime.shadow_engine", shadow)
    monkeypatch.setitem(sys.modules, "atomo_b200.runtime.engine", fused)

    def args(*argv):
        return add_fit_args(argparse.ArgumentParser(), list(argv))

    _, kind = L._build_engine(args("--dtype", "bf16", "--code", "svd", "--svd-rank", "3", "--num-aggregate", "2"), None, 0, 4)
    assert kind == "shadow" and made[-1][0] == "ShadowEngine"
    assert made[-1][1]["ps_mode"] == "sharded" and made[-1][1]["num_aggregate"] == 2 and made[-1][1]["groups"] == 5
    _, kind = L._build_engine(args("--dtype", "bf16", "--code", "qsvd", "--optimizer", "adam"), None, 0, 2)
    assert kind == "shadow" and made[-1][1]["optimizer"] == "adam" and made[-1][1]["code"] == "qsvd"
    _, kind = L._build_engine(args("--dtype", "bf16", "--code", "qsgd"), None, 0, 2)
    assert kind == "fused" and made[-1][0] == "FusedEngine" and made[-1][1]["ps_mode"] == "colocated"
    _, kind = L._build_engine(args("--dtype", "fp32", "--code", "svd", "--ps-mode", "dedicated"), None, 0, 2)
    assert kind == "fused" and made[-1][1]["ps_mode"] == "dedicated" and made[-1][1]["dtype"] == "fp32"
    import pytest
    with pytest.raises(SystemExit):
        L._build_engine(args("--dtype", "fp32", "--code", "svd", "--optimizer", "adam"), None, 0, 2)


def test_worker_log_line_roundtrips_through_the_tuning_parser():
    """The worker line is a de-facto API (the LR grid search greps it): whatever the trainer can print — diverged
    losses included — the parser must read back."""
    from hypothesis import given, settings, strategies as st
    from atomo_b200.tiny_tuning_parser import parse_line
    from atomo_b200.utils.logging import worker_line

    pos = st.floats(0, 1e4, allow_nan=False)

    @settings(max_examples=200, deadline=None)
    @given(rank=st.integers(0, 64), step=st.integers(0, 10 ** 7), epoch=st.integers(0, 999), seen=st.integers(0, 10 ** 6),
           total=st.integers(1, 10 ** 7), loss=st.one_of(st.floats(-1e3, 1e6, allow_nan=False), st.just(float("nan")),
                                                           st.just(float("inf"))),
           t=pos, comp=pos, enc=pos, comm=pos, mb=pos, p1=st.floats(0, 100), p5=st.floats(0, 100))
    def check(rank, step, epoch, seen, total, loss, t, comp, enc, comm, mb, p1, p5):
        rec = parse_line(worker_line(rank, step, epoch, seen, total, loss, t, comp, enc, comm, mb, p1, p5))
        assert rec is not None and rec["worker"] == rank and rec["step"] == step
        if loss == loss and abs(loss) != float("inf"):
            assert abs(rec["loss"] - loss) <= 5e-5 + 1e-9 * abs(loss)
        else:
            assert rec["loss"] != rec["loss"] or abs(rec["loss"]) == float("inf")
        assert abs(rec["msg_mb"] - mb) <= 5e-5 and abs(rec["prec1"] - p1) <= 5e-5

    check()
