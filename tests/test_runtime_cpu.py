"""PS/worker plumbing on CPU: wire format, plan tables, gloo multi-process runs, checkpoints, evaluator, parser."""
import os
import struct
import subprocess
import sys

import pytest
import torch

from atomo_b200 import codings
from atomo_b200.ops import plan as P
from atomo_b200.parallel import wire
from atomo_b200.utils import checkpoint as ckpt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wire_roundtrip_all_coders():
    g = torch.randn(12, 6, 3, 3)
    msgs = [codings.build(n, **kw).encode(g) for n, kw in
            [("svd", dict(rank=2)), ("qsgd", {}), ("entrywise", dict(budget=0.2)), ("sgd", {}),
             ("sgd", dict(compress=True)), ("qsvd", dict(rank=2))]]
    buf = wire.pack({"step": 9, "rank": 1, "codes": msgs})
    assert buf.dtype == torch.uint8
    out = wire.unpack(buf)
    assert out["step"] == 9 and len(out["codes"]) == len(msgs)
    for name, a, b in zip(["svd", "qsgd", "entrywise", "sgd", "sgd", "qsvd"], msgs, out["codes"]):
        coder = codings.build(name, rank=2) if name in ("svd", "qsvd") else codings.build(name)
        assert torch.allclose(coder.decode(a), coder.decode(b))
    with pytest.raises(ValueError):
        wire.unpack(torch.zeros(64, dtype=torch.uint8))


def test_plan_tables_resnet18():
    from atomo_b200.models import build_model
    from atomo_b200.runtime.flat import FlatLayout
    m = build_model("ResNet18")
    lay = FlatLayout.from_module(m)
    pl = P.build_plan(lay.shapes, "svd", 3, offsets=lay.offsets)
    assert pl.total_elems == lay.total
    ts = [l for l in pl.layers if l.route == P.ROUTE_SVD_TS]
    assert len(ts) == 18  # 17 3x3 convs + fc (transposed); the square-ish 1x1 shortcuts travel dense
    conv = next(l for l in ts if l.shape == (512, 512, 3, 3))
    assert (conv.rows, conv.cols, conv.row_stride, conv.col_stride) == (131072, 18, 18, 1)
    fc = next(l for l in ts if l.shape == (10, 512))
    assert (fc.rows, fc.cols, fc.row_stride, fc.col_stride) == (512, 10, 1, 512)
    assert all(l.route == P.ROUTE_DENSE for l in pl.layers if len(l.shape) == 1)
    # tiles cover every row / element exactly once
    for l in ts:
        rows = sum(t[2] for t in pl.enc_tiles if t[0] == l.index)
        assert rows == l.rows
        rows = sum(t[2] for t in pl.ps_tiles if t[0] == l.index)
        assert rows == l.rows
        assert l.ps_rows * l.cols <= P.PS_TILE_ELEMS and l.ps_rows * ((l.cols + 3) // 4) <= 1024
        assert l.slot_off % 32 == 0 and P.slot_u_off(l.rcap, l.cols) % 4 == 0
    for l in pl.layers:
        if l.route == P.ROUTE_DENSE:
            assert sum(t[2] for t in pl.ps_tiles if t[0] == l.index) == l.numel
    assert len(pl.layers_bytes()) == 72 * 62 and P.CTRL_BYTES == 56
    step, err, lr = struct.unpack_from("<iif", P.pack_ctrl(step=4, lr=0.5))
    assert (step, err) == (4, 0) and lr == 0.5
    # r=3 factors are several times smaller than the dense gradient (~6x at the reference's sizes)
    assert pl.factor_bytes_per_worker() < 0.5 * 4 * lay.total


def test_slot_capacity_rules():
    assert P.slot_capacity(18, 3, False) == 8 and P.slot_capacity(18, 3, True) == 4
    assert P.slot_capacity(2, 3, False) == 4 and P.slot_capacity(50, 16, False) == 32
    assert P.slot_capacity(18, 0, False) == 20


def test_checkpoint_layout_and_resume(tmp_path):
    net = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    d = str(tmp_path) + "/"
    path = ckpt.save_model(d, 50, net)
    assert path == d + "model_step_50" and os.path.isfile(path)  # master:331-336 naming
    ckpt.save_sidecar(d, 50, opt, lr=0.05, extra={"shrink_counter": 1})
    ckpt.save_model(d, 100, net)
    assert ckpt.latest_step(d) == 100
    net2 = torch.nn.Linear(4, 2)
    ckpt.load_model(d, 50, net2)
    assert torch.equal(net2.weight, net.weight)
    side = ckpt.load_sidecar(d, 50, opt)
    assert side["lr"] == 0.05 and side["step"] == 50 and side["shrink_counter"] == 1
    assert ckpt.load_sidecar(d, 100) is None


def _run_launcher(extra, timeout=480):
    cmd = [sys.executable, "-m", "atomo_b200.distributed_nn", "--synthetic", "1", "--train-len", "512",
           "--test-len", "128", "--batch-size", "32", "--lr", "0.05", "--test-batch-size", "64"] + extra
    # 2-3 ranks x (8 OpenMP threads) oversubscribe the 8-core CI box: pin the ranks to 2 threads each
    env = dict(os.environ, ATOMO_HANG_DUMP_S=str(timeout - 20), PYTHONPATH=ROOT, OMP_NUM_THREADS="2",
               MKL_NUM_THREADS="2")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_baseline_config1_lenet_entrywise_gloo_world2(tmp_path):
    """BASELINE.json config 1: LeNet, MNIST-shaped synthetic, entry-wise sparsifier, PS + 1 worker on CPU/gloo."""
    from atomo_b200.tiny_tuning_parser import parse_line
    d = str(tmp_path) + "/"
    out = _run_launcher(["--nproc", "2", "--network", "LeNet", "--dataset", "MNIST", "--code", "entrywise",
                         "--entry-budget", "0.1", "--max-steps", "8", "--eval-freq", "4", "--train-dir", d,
                         "--master-port", "29541"])
    recs = [parse_line(l) for l in out.splitlines() if l.startswith("Worker:")]
    assert len(recs) == 8 and all(r is not None for r in recs)
    assert recs[-1]["loss"] < recs[0]["loss"]
    assert recs[0]["msg_mb"] < 0.5  # 10% of 1.64 MiB (x2 for indices)
    assert "Master: Step: 8" in out and "Test set: Step: 4" in out
    assert os.path.isfile(d + "model_step_4") and os.path.isfile(d + "model_step_8_optim")
    # the polling evaluator consumes the same files (distributed_evaluator.py contract)
    from atomo_b200.distributed_evaluator import main as ev_main
    res = ev_main(["--model-dir", d, "--eval-freq", "4", "--network", "LeNet", "--dataset", "MNIST", "--synthetic", "1",
                   "--test-len", "128", "--eval-batch-size", "64", "--max-evals", "2", "--poll-seconds", "0.1"])
    assert [r["step"] for r in res] == [4, 8]


@pytest.mark.parametrize("code,extra", [("svd", ["--svd-rank", "3"]), ("qsgd", ["--quantization-level", "4"]),
                                        ("sgd", []), ("qsvd", ["--svd-rank", "2"]), ("bsvd", ["--svd-rank", "4"])])
def test_three_rank_gloo_all_coders(code, extra, tmp_path):
    m = str(tmp_path / "metrics")
    out = _run_launcher(["--nproc", "3", "--network", "LeNet", "--dataset", "MNIST", "--code", code,
                         "--max-steps", "4", "--eval-freq", "100", "--train-dir", str(tmp_path) + "/",
                         "--master-port", str(29550 + len(code) + len(extra)), "--metrics-file", m] + extra)
    assert out.count("Worker: 1, Step:") == 4 and out.count("Worker: 2, Step:") == 4
    assert "Master: Step: 4" in out
    # --metrics-file mirrors the log lines, one JSON object per logged step and rank
    import json
    from atomo_b200.tiny_tuning_parser import parse_line
    ps = [json.loads(l) for l in open(m + ".rank0.jsonl")]
    w1 = [json.loads(l) for l in open(m + ".rank1.jsonl")]
    assert [r["step"] for r in ps] == [1, 2, 3, 4] and all(r["role"] == "ps" and r["used_workers"] == [1, 2] for r in ps)
    assert [r["step"] for r in w1] == [1, 2, 3, 4] and all(r["role"] == "worker" and r["msg_mb"] > 0 for r in w1)
    line = [parse_line(l) for l in out.splitlines() if l.startswith("Worker: 1, Step: 4,")][0]
    assert abs(line["loss"] - w1[-1]["loss"]) < 1e-3


def test_resume_continues_from_checkpoint(tmp_path):
    d = str(tmp_path) + "/"
    common = ["--nproc", "2", "--network", "LeNet", "--dataset", "MNIST", "--code", "sgd", "--eval-freq", "3",
              "--train-dir", d]
    _run_launcher(common + ["--max-steps", "3", "--master-port", "29581"])
    out = _run_launcher(common + ["--max-steps", "5", "--resume", "1", "--master-port", "29582"])
    assert "Master: Step: 4" in out and "Master: Step: 3," not in out


def test_single_machine_and_tuning_parser(tmp_path):
    from atomo_b200.single_machine import main as sm_main
    res = sm_main(["--network", "LeNet", "--dataset", "MNIST", "--synthetic", "1", "--train-len", "256", "--test-len",
                   "64", "--batch-size", "32", "--max-steps", "6", "--test-batch-size", "64", "--lr", "0.05"])
    assert res["loss"] > 0
    from atomo_b200.tiny_tuning_parser import main as tp_main
    from atomo_b200.utils.logging import worker_line
    f = tmp_path / "0.01"
    f.write_text("\n".join(worker_line(w, 100, 0, 0, 100, 1.0 + w, 0.1, 0.1, 0.1, 0.1, 1.0, 10, 50) for w in (1, 2, 3)))
    assert tp_main(["--tuning-dir", str(tmp_path), "--tuning-lr", "0.01", "--num-workers", "3"]) == pytest.approx(3.0)


def test_single_machine_with_a_coder_in_the_loop():
    from atomo_b200.single_machine import main as sm_main
    for code in ("bsvd", "entrywise"):
        res = sm_main(["--network", "LeNet", "--dataset", "MNIST", "--synthetic", "1", "--train-len", "256",
                       "--test-len", "64", "--batch-size", "32", "--max-steps", "8", "--test-batch-size", "64",
                       "--lr", "0.02", "--code", code, "--svd-rank", "3", "--entry-budget", "0.2"])
        assert 0 < res["loss"] < 2.4


def test_data_sharding_and_loader():
    from atomo_b200.data import DataLoader, build_datasets, shard_indices
    a, b = shard_indices(100, 0, 2, seed=3), shard_indices(100, 1, 2, seed=3)
    assert len(set(a.tolist()) & set(b.tolist())) == 0 and len(a) == len(b) == 50
    tr, te, nc = build_datasets("Cifar100", synthetic=True, train_len=64, test_len=16)
    assert nc == 100 and tr[0][0].shape == (3, 32, 32)
    ld = DataLoader(tr, batch_size=16, shuffle=True, prefetch=2)
    for _ in range(6):  # wraps over the epoch boundary (persistent iterator, my_data_loader.py:310-319)
        x, y = ld.next_batch()
        assert x.shape == (16, 3, 32, 32)
    assert ld.epochs_completed >= 1
    ld.close()


def test_straggler_kill_split_backward(tmp_path):
    """--num-aggregate 1 of 2 workers + --straggler-kill: workers run the layer-wise (split) backward with a
    kill listener, the PS signals the straggler and drops whatever it still sends; training completes."""
    out = _run_launcher(["--nproc", "3", "--network", "LeNet", "--dataset", "MNIST", "--code", "svd", "--svd-rank", "2",
                         "--max-steps", "5", "--num-aggregate", "1", "--straggler-kill", "1", "--eval-freq", "100",
                         "--train-dir", str(tmp_path) + "/", "--master-port", "29591"])
    assert "Master: Step: 5" in out
    steps = [l for l in out.splitlines() if l.startswith("Worker:")]
    assert len(steps) >= 5  # every step is completed by at least one worker (others may have been abandoned)


def test_kill_signal_is_step_stamped():
    from atomo_b200.parallel.transport import TorchDistTransport

    class KV(dict):
        def check(self, keys): return all(k in self for k in keys)
        def get(self, k): return self[k].encode()
        def set(self, k, v): self[k] = v

    t = TorchDistTransport.__new__(TorchDistTransport)
    t.rank, t._kv, t._kill_enabled = 2, KV(), True
    assert not t.kill_requested(3)
    t.send_kill(2, 7)
    assert not t.kill_requested(8) and t.kill_requested(7)   # a step-7 signal cannot abort step 8


def test_ps_checkpoint_of_a_bn_network_carries_trained_statistics(tmp_path):
    """the PS never runs a forward pass, so its own BatchNorm buffers stay at (0, 1).  The first worker
    ships its running statistics with the gradients of every checkpoint step; the saved model_step_<N> must
    therefore hold trained buffers (the evaluator loads it under .eval())."""
    d = str(tmp_path) + "/"
    _run_launcher(["--nproc", "2", "--network", "ResNet18", "--dataset", "Cifar10", "--code", "sgd", "--batch-size", "8",
                   "--max-steps", "2", "--eval-freq", "2", "--train-dir", d, "--master-port", "29597"], timeout=400)
    sd = torch.load(d + "model_step_2", weights_only=False)
    assert float(sd["bn1.running_mean"].abs().sum()) > 0
    assert not torch.allclose(sd["bn1.running_var"], torch.ones_like(sd["bn1.running_var"]))
    assert int(sd["bn1.num_batches_tracked"]) >= 2


def _run_backup(extra, env_extra, port, nproc=3, steps=12, need=1, code=("--code", "sgd")):
    cmd = [sys.executable, "-m", "atomo_b200.distributed_nn", "--synthetic", "1", "--train-len", "512",
           "--test-len", "128", "--batch-size", "32", "--lr", "0.05", "--test-batch-size", "64",
           "--nproc", str(nproc), "--network", "LeNet", "--dataset", "MNIST", *code, "--max-steps", str(steps),
           "--num-aggregate", str(need), "--master-port", str(port)] + extra
    env = dict(os.environ, ATOMO_HANG_DUMP_S="440", PYTHONPATH=ROOT, OMP_NUM_THREADS="2", MKL_NUM_THREADS="2",
               **env_extra)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=480)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_backup_workers_do_not_wait_for_a_straggler(tmp_path):
    """with --num-aggregate 1 of 2 the PS used to drain every straggler message before the next step,
    so each step still cost the slowest worker's time.  Rounds are now announced point-to-point to the workers that
    owe nothing (parallel/backup_rounds.py): a worker that sleeps 0.5 s per step must not slow the PS down, its late
    gradients are dropped, and it rejoins with the CURRENT parameters (so it logs far fewer steps than the fast
    worker, and the steps it does log are not consecutive replays)."""
    import re
    out = _run_backup(["--eval-freq", "100", "--train-dir", str(tmp_path) + "/"],
                      {"ATOMO_DEBUG_SLOW_WORKER": "2:0.5"}, 29597)
    assert "Master: Step: 12" in out
    gathers = [float(m.group(1)) for m in re.finditer(r"Master: Step: \d+, .*Gather: ([0-9.e-]+)", out)]
    assert len(gathers) == 12
    # never the straggler's 0.5 s per step (the first two steps are excluded: process start-up on a loaded box can
    # make the "fast" worker the slower one once)
    assert sum(gathers[2:]) < 0.5 * 10 * 0.5, gathers
    fast = [int(m.group(1)) for m in re.finditer(r"Worker: 1, Step: (\d+),", out)]
    slow = [int(m.group(1)) for m in re.finditer(r"Worker: 2, Step: (\d+),", out)]
    # the fast worker is handed the current step the moment its previous message lands: it takes part in (almost)
    # every step, the straggler in a few, and never in consecutive replays of what it missed
    assert len(fast) >= 10 and fast[-1] == 12 and fast == sorted(set(fast)), fast
    assert 1 <= len(slow) < 6, slow


def test_ps_survives_lost_workers_in_backup_mode(tmp_path):
    """SURVEY 5.3 (the reference's PS blocks forever in waitany when a worker dies): with --num-aggregate 2 of 3,
    worker 2 dies in step 3 and worker 3 in step 5.  The PS notices each closed connection, keeps training with the
    survivors (need shrinks to 1), writes its checkpoints and shuts down cleanly."""
    d = str(tmp_path) + "/"
    out = _run_backup(["--eval-freq", "4", "--train-dir", d], {"ATOMO_DEBUG_DIE_WORKER": "2:3,3:5"}, 29598,
                      nproc=4, steps=8, need=2, code=("--code", "svd", "--svd-rank", "2"))
    assert "Master: worker 2 is gone" in out and "Master: worker 3 is gone" in out
    assert "Master: Step: 8" in out
    # worker 1 may miss the cut of an early step (2 of 3 are enough) but it carries steps 6..8 alone
    assert all("Worker: 1, Step: %d," % s in out for s in (6, 7, 8))
    assert "Done sending messages to workers!" in out
    assert os.path.isfile(d + "model_step_8") and os.path.isfile(d + "model_step_8_optim")


def test_ps_fails_fast_when_a_worker_dies_in_all_workers_mode(tmp_path):
    """Without backup workers the PS needs every gradient: a dead worker must stop the job at once with a clear
    message (the reference blocks forever in waitany; a plain gloo receive would wait for the 30-minute timeout)."""
    import time
    cmd = [sys.executable, "-m", "atomo_b200.distributed_nn", "--synthetic", "1", "--train-len", "512",
           "--test-len", "128", "--batch-size", "32", "--lr", "0.05", "--test-batch-size", "64", "--nproc", "3",
           "--network", "LeNet", "--dataset", "MNIST", "--code", "sgd", "--max-steps", "50", "--eval-freq", "100",
           "--train-dir", str(tmp_path) + "/", "--master-port", "29599"]
    env = dict(os.environ, ATOMO_HANG_DUMP_S="440", PYTHONPATH=ROOT, OMP_NUM_THREADS="2", MKL_NUM_THREADS="2",
               ATOMO_DEBUG_DIE_WORKER="2:3")
    t0 = time.time()
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=480)
    assert r.returncode != 0
    assert "Master: worker 2 is gone" in r.stdout and "stopping the job" in r.stdout
    assert "Master: Step: 2," in r.stdout and "Master: Step: 4," not in r.stdout
    assert time.time() - t0 < 400          # not the 30-minute process-group timeout


def test_launcher_restarts_from_the_latest_checkpoint(tmp_path):
    """--max-restarts: worker 2 dies in step 5 (once), the PS stops the job, the launcher relaunches it with --resume
    and training continues from the step-4 checkpoint to the end."""
    d = str(tmp_path) + "/"
    cmd = [sys.executable, "-m", "atomo_b200.distributed_nn", "--synthetic", "1", "--train-len", "512",
           "--test-len", "128", "--batch-size", "32", "--lr", "0.05", "--test-batch-size", "64", "--nproc", "3",
           "--network", "LeNet", "--dataset", "MNIST", "--code", "svd", "--svd-rank", "2", "--max-steps", "8",
           "--eval-freq", "2", "--train-dir", d, "--master-port", "29601", "--max-restarts", "2"]
    env = dict(os.environ, ATOMO_HANG_DUMP_S="440", PYTHONPATH=ROOT, OMP_NUM_THREADS="2", MKL_NUM_THREADS="2",
               ATOMO_DEBUG_DIE_WORKER="2:5", ATOMO_DEBUG_DIE_ONCE=str(tmp_path / "died"))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=480)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "Master: worker 2 is gone" in out and "restart 1/2 from the latest checkpoint" in out
    assert out.count("Master: Step: 4,") == 1          # not replayed: the relaunch starts at step 5
    assert out.count("Master: Step: 5,") == 1 and "Master: Step: 8," in out
    assert os.path.isfile(d + "model_step_8")
