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


# ---------------------------------------------------------------------------------------------------------
# v2: backward groups pushed while backward continues, sharded owners, optional backup-worker aggregation
# ---------------------------------------------------------------------------------------------------------
def run_model_v2(W=4, G=3, owners=None, steps=25, seed=0, need=0, slow_worker=None):
    """Executable model of runtime/shadow_engine.py + csrc/v2_*.cu.

    Every rank is a worker; ``owners`` are the ranks that host PS shards (all ranks = sharded PS).  Per step a
    worker walks the groups in backward order; group g's "weights" may only be read before its push.  Each
    owner thread serves its shard of every group: waits for the push flags [g][w] == step (all W, or the first
    ``need`` with a published mask), checks the step stamped in every slot it consumes, updates ITS shard of the
    group's weights on every rank, and after the last group publishes param_flag[owner] = step + 1.
    Invariants: no deadlock; a worker never reads a weight of step t that was already overwritten (or not yet
    written); an owner never consumes a slot of another step; all ranks end with identical weights."""
    rnd = random.Random(seed)
    owners = list(range(W)) if owners is None else owners
    nO = len(owners)
    # weights[rank][g][shard]: value = number of updates applied
    weights = [[[0] * nO for _ in range(G)] for _ in range(W)]
    param_flag = [[1] * nO for _ in range(W)]              # on every rank: one flag per owner
    push_flag = [[[0] * W for _ in range(G)] for _ in range(nO)]   # on every owner: [g][w]
    slots = [[[None] * W for _ in range(G)] for _ in range(nO)]    # on every owner: [g][w] -> (step, value)
    errors, done = [], threading.Event()
    lock = threading.Lock()
    owner_step = [1] * nO

    def jitter(scale=1.0):
        time.sleep(rnd.random() * 0.0015 * scale)

    def worker(w):
        t = 1
        while t <= steps and not done.is_set():
            # wait_params: every owner has published the parameters of step t (stragglers may skip ahead)
            while min(param_flag[w]) < t:
                if done.is_set():
                    return
                time.sleep(0)
            if need and min(param_flag[w]) > t:
                t = min(param_flag[w])                     # skip-ahead of a straggler that was left out
                if t > steps:
                    return
            # forward reads every weight: all must be exactly the step-(t-1) values
            snap = [list(weights[w][g]) for g in range(G)]
            if not need and any(v != t - 1 for row in snap for v in row):
                errors.append(("forward saw torn/stale weights", w, t, snap))
            for g in range(G):                             # backward, group by group
                jitter(4.0 if w == slow_worker else 1.0)
                # backward of group g still reads group g's weights: they must not have moved yet
                if not need and any(v != t - 1 for v in weights[w][g]):
                    errors.append(("weights of a group changed before its backward finished", w, t, g))
                for o in range(nO):                        # project: factors into every owner's arena, then flag
                    slots[o][g][w] = (t, 1)
                jitter(0.3)
                for o in range(nO):
                    push_flag[o][g][w] = t                 # st.release.sys; flags only increase
            t += 1

    def owner(o):
        for t in range(1, steps + 1):
            for g in range(G):
                # the PS launch of group g is stream-ordered after this rank's own push of group g
                while push_flag[o][g][owners[o]] < t and not need:
                    if done.is_set():
                        return
                    time.sleep(0)
                want = need if need else W
                while True:
                    ready = [w for w in range(W) if push_flag[o][g][w] >= t]
                    if len(ready) >= want:
                        break
                    if done.is_set():
                        return
                    time.sleep(0)
                ready = ready[:want] if need else ready
                acc = 0
                for w in ready:
                    st, val = slots[o][g][w]
                    if st != t:
                        errors.append(("owner consumed a slot of another step", o, t, g, w, st))
                    acc += val
                jitter(0.5)
                new = t                                    # "apply the averaged gradient": weights become t
                for r in range(W):                         # multicast of this owner's shard of group g
                    weights[r][g][o] = new
            for r in range(W):
                param_flag[r][o] = t + 1                   # after the LAST group: parameters of t+1 in place
            owner_step[o] = t + 1

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)] + \
              [threading.Thread(target=owner, args=(o,)) for o in range(nO)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=90)
    alive = any(th.is_alive() for th in threads)
    done.set()
    same = all(weights[r] == weights[0] for r in range(W))
    return errors, alive, same, weights[0][0][0]


def test_v2_protocol_sharded_and_centralized():
    for owners in (None, [0]):                 # sharded PS, centralized (colocated) PS
        for seed in range(2):
            errors, alive, same, final = run_model_v2(W=4, G=3, owners=owners, steps=20, seed=seed)
            assert not alive, "deadlock"
            assert not errors, errors[:3]
            assert same and final == 20


def test_v2_protocol_backup_workers_with_a_straggler():
    """--num-aggregate 3 of 4 with one worker 4x slower: the owner proceeds without it, never consumes a slot of
    another step, and nobody deadlocks (the straggler skips ahead to the published step)."""
    errors, alive, same, final = run_model_v2(W=4, G=3, owners=[0], steps=15, seed=1, need=3, slow_worker=3)
    assert not alive, "deadlock"
    assert not [e for e in errors if e[0] == "owner consumed a slot of another step"], errors[:3]
    assert same and final == 15
