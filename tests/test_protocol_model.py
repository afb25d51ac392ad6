"""Executable model of the step-stamped flag protocol of the fused engine, run under
randomised thread delays: no deadlock, no torn parameter read, no stale slot consumed, flags never reset.

Memory-model effects (release/acquire over NVLink) are covered on hardware by
tests/test_gpu_engine.py::test_two_gpu_engine_keeps_replicas_identical; this model checks the *logic*:
who may write what when, given only monotonically increasing flags.
"""
import random
import threading
import time


class Rank:
    def __init__(self, nparams):
        self.params = [0] * nparams      # parameter buffer in this rank's HBM (written by the PS)
        self.param_flag = 1              # "parameters of step t are in place"
        self.step = 1                    # device-side step counter (Ctrl::step)


def run_model(W=4, steps=40, nparams=16, seed=0, colocated=True):
    rnd = random.Random(seed)
    ranks = [Rank(nparams) for _ in range(W)]
    push_flags = [0] * W                 # on the PS
    slots = [None] * W                   # on the PS: (step, payload) written by worker w
    ps_params = [0] * nparams
    errors, done = [], threading.Event()

    def jitter():
        time.sleep(rnd.random() * 0.002)

    def worker(w):
        r = ranks[w]
        for _ in range(steps):
            t = r.step
            while r.param_flag < t:               # wait_params_kernel
                if done.is_set():
                    return
                time.sleep(0)
            snapshot = list(r.params)             # forward/backward reads the parameters
            jitter()
            if any(v != snapshot[0] for v in snapshot) or snapshot[0] != t - 1:
                errors.append(("torn or stale params", w, t, snapshot[:4]))
            slots[w] = (t, [snapshot[0] + 1] * nparams)   # encode + peer stores into the PS arena
            jitter()
            push_flags[w] = t                     # st.release.sys (flags only ever increase)
            if w == 0 and colocated:
                ps_update(t)
            r.step = t + 1                        # advance_step_kernel

    def ps_update(t):
        for w in range(W):                        # spin on every worker's push flag
            while push_flags[w] < t:
                if done.is_set():
                    return
                time.sleep(0)
        acc = 0
        for w in range(W):
            st, payload = slots[w]
            if st != t:
                errors.append(("stale slot", w, t, st))
            acc += payload[0]
        new = acc // W                            # averaged "gradient" applied: params become t
        for i in range(nparams):
            ps_params[i] = new
        for r in ranks:                           # multicast store of the parameter tiles ...
            for i in range(nparams):
                r.params[i] = new
                if i == nparams // 2:
                    jitter()
        for r in ranks:                           # ... then the flag
            r.param_flag = t + 1

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    alive = any(th.is_alive() for th in threads)
    done.set()
    return errors, alive, ps_params[0]


def test_flag_protocol_has_no_deadlock_or_torn_reads():
    for seed in range(3):
        errors, alive, final = run_model(W=4, steps=30, seed=seed)
        assert not alive, "deadlock"
        assert not errors, errors[:3]
        assert final == 30
