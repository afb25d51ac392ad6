"""Exhaustive interleaving check of the grouped / sharded flag protocol (runtime/shadow_engine.py + csrc/v2_*.cu).

tests/test_protocol_model.py runs the protocol under random thread delays; this file explores EVERY interleaving of a
small configuration (breadth-first over the global state space, one shared-memory access per atomic step) and checks

* no deadlock (some process can always move until all have finished),
* a worker never reads a weight that is not exactly the previous step's value (torn / early / stale parameters),
* an owner never consumes a slot stamped with another step,
* all replicas end identical.

Sequential consistency is assumed (release/acquire over NVLink is what the hardware tests cover); what is verified
here is the LOGIC: which flag guards which buffer.  Three deliberately broken variants must be caught, so the checker
is known to have teeth.
"""
from collections import deque

import pytest


def build(W, G, owners, T, bug=None):
    """Processes as straight-line programs of atomic instructions over a flat memory dict.
    Instruction kinds: ('wait', fn(mem)->bool), ('set', key, value), ('check', fn(mem)->bool, message)."""
    nO = len(owners)
    mem = {}
    for r in range(W):
        for g in range(G):
            for o in range(nO):
                mem[("w", r, g, o)] = 0                  # weight shard (rank r, group g, owner o): #updates applied
        for o in range(nO):
            mem[("pf", r, o)] = 1                        # param flag of owner o on rank r
    for o in range(nO):
        for g in range(G):
            for w in range(W):
                mem[("push", o, g, w)] = 0
                mem[("slot", o, g, w)] = 0
    procs = []
    for w in range(W):
        prog = []
        for t in range(1, T + 1):
            prog.append(("wait", lambda m, w=w, t=t: all(m[("pf", w, o)] >= t for o in range(nO))))
            for g in range(G):                            # forward reads every weight
                for o in range(nO):
                    prog.append(("check", lambda m, w=w, g=g, o=o, t=t: m[("w", w, g, o)] == t - 1,
                                 "forward of step %d read a weight that is not of step %d" % (t, t - 1)))
            for g in range(G):                            # backward, group by group
                for o in range(nO):
                    prog.append(("check", lambda m, w=w, g=g, o=o, t=t: m[("w", w, g, o)] == t - 1,
                                 "weights of a group moved before its backward finished"))
                stores = [("set", ("slot", o, g, w), t) for o in range(nO)]
                flags = [("set", ("push", o, g, w), t) for o in range(nO)]
                prog += (flags + stores) if bug == "flag_before_slot" else (stores + flags)
        procs.append(prog)
    for oi, orank in enumerate(owners):
        prog = []
        for t in range(1, T + 1):
            for g in range(G):
                if orank < W and bug != "no_stream_order":
                    # the PS launch of group g is stream-ordered after this rank's own push of group g
                    prog.append(("wait", lambda m, oi=oi, g=g, t=t, me=orank: m[("push", oi, g, me)] >= t))
                prog.append(("wait", lambda m, oi=oi, g=g, t=t: all(m[("push", oi, g, w)] >= t for w in range(W))))
                for w in range(W):
                    prog.append(("check", lambda m, oi=oi, g=g, w=w, t=t: m[("slot", oi, g, w)] == t,
                                 "owner consumed a slot of another step"))
                writes = [("set", ("w", r, g, oi), t) for r in range(W)]
                if bug == "publish_early" and g == G - 1:
                    prog += [("set", ("pf", r, oi), t + 1) for r in range(W)] + writes
                else:
                    prog += writes
            if bug != "publish_early":
                prog += [("set", ("pf", r, oi), t + 1) for r in range(W)]
        procs.append(prog)
    return mem, procs


def explore(W, G, owners, T, bug=None, limit=3_000_000):
    mem0, procs = build(W, G, owners, T, bug)
    keys = sorted(mem0)
    kidx = {k: i for i, k in enumerate(keys)}
    start = (tuple(0 for _ in procs), tuple(mem0[k] for k in keys))
    seen = {start}
    todo = deque([start])
    ends = set()
    while todo:
        pcs, vals = todo.popleft()
        mem = None
        moved = False
        for p, prog in enumerate(procs):
            pc = pcs[p]
            if pc >= len(prog):
                continue
            ins = prog[pc]
            if mem is None:
                mem = dict(zip(keys, vals))
            if ins[0] == "wait":
                if not ins[1](mem):
                    continue
                nvals = vals
            elif ins[0] == "check":
                if not ins[1](mem):
                    return {"error": ins[2], "states": len(seen)}
                nvals = vals
            else:
                i = kidx[ins[1]]
                nvals = vals[:i] + (ins[2],) + vals[i + 1:]
            moved = True
            nxt = (pcs[:p] + (pc + 1,) + pcs[p + 1:], nvals)
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
                if len(seen) > limit:
                    return {"error": "state space larger than %d" % limit, "states": len(seen)}
        if not moved:
            if any(pc < len(prog) for pc, prog in zip(pcs, procs)):
                return {"error": "deadlock", "states": len(seen)}
            ends.add(vals)
    final = [dict(zip(keys, v)) for v in ends]
    same = all(all(m[("w", r, g, o)] == T for r in range(W) for g in range(G) for o in range(len(owners))) for m in final)
    return {"error": None, "states": len(seen), "replicas_identical": same}


@pytest.mark.parametrize("W,G,owners,T", [
    (2, 2, [0, 1], 2),      # sharded PS, two backward groups, two steps (flag reuse across steps)
    (2, 2, [0], 2),         # centralized, colocated PS
    (2, 1, [2], 2),         # dedicated PS: the owner is not a worker
    (3, 1, [0, 1, 2], 2),   # three workers / three owners, one group
    (3, 2, [0], 2),         # three workers, centralized PS, two groups
    (2, 3, [0, 1], 3),      # three groups, three steps: every flag is reused twice (52 k states)
    (4, 1, [0, 1, 2, 3], 2),  # four workers / four owners (0.8 M states)
])
def test_every_interleaving_is_safe(W, G, owners, T):
    res = explore(W, G, owners, T)
    assert res["error"] is None, res
    assert res["replicas_identical"] and res["states"] >= 90


@pytest.mark.parametrize("bug,expect", [
    ("flag_before_slot", "owner consumed a slot of another step"),
    ("publish_early", "read a weight that is not of step"),
])
def test_the_checker_catches_broken_protocols(bug, expect):
    res = explore(2, 2, [0, 1], 2, bug=bug)
    assert res["error"] and expect in res["error"], res


def test_stream_order_between_own_push_and_ps_launch_is_not_load_bearing():
    """The PS launch of a group is stream-ordered after the same rank's push of that group; the wait for ALL workers'
    flags subsumes it, so the protocol stays safe without it (it is an optimisation: no spinning CTAs before the data
    can possibly be there)."""
    res = explore(2, 2, [0, 1], 2, bug="no_stream_order")
    assert res["error"] is None and res["replicas_identical"]


# ---------------------------------------------------------------------------------------------------------
# backup workers (--num-aggregate N < W): data-dependent control flow (skip-ahead, aggregation mask), so the
# processes are small state machines instead of straight-line programs
# ---------------------------------------------------------------------------------------------------------
def explore_backup(W, G, owners, T, need, limit=3_000_000, bug=None):
    """Owner: waits until `need` workers pushed group g of its step, picks ANY `need` of the ready ones (every choice
    is explored), checks their slot stamps, updates, publishes.  Worker: waits for every owner's parameter flag,
    skips ahead to min(flag) (a straggler that was left out does not replay old steps), pushes every group to every
    owner.  Weights may be read while they change (inherent to in-place updates with backup workers), so only
    deadlock-freedom, slot stamps and 'a worker is never ahead of an owner' are checked."""
    from itertools import combinations
    nO = len(owners)
    keys, mem0 = [], {}
    for r in range(W):
        for o in range(nO):
            mem0[("pf", r, o)] = 1
    for o in range(nO):
        for g in range(G):
            for w in range(W):
                mem0[("push", o, g, w)] = 0
                mem0[("slot", o, g, w)] = 0
    keys = sorted(mem0)
    kidx = {k: i for i, k in enumerate(keys)}

    def get(vals, k):
        return vals[kidx[k]]

    def put(vals, k, v):
        i = kidx[k]
        return vals[:i] + (v,) + vals[i + 1:]

    def worker_moves(w, loc, vals):
        t, phase, g, o = loc
        if phase == "done":
            return None
        if phase == "wait":
            flags = [get(vals, ("pf", w, x)) for x in range(nO)]
            if min(flags) < t:
                return []
            t2 = t if bug == "no_skip_ahead" else min(flags)
            return [((t2, "done", 0, 0), vals, None)] if t2 > T else [((t2, "slot", 0, 0), vals, None)]
        if phase == "slot":
            nv = put(vals, ("slot", o, g, w), t)
            return [((t, "slot", g, o + 1) if o + 1 < nO else (t, "flag", g, 0), nv, None)]
        nv = put(vals, ("push", o, g, w), t)                      # phase == "flag"
        if o + 1 < nO:
            return [((t, "flag", g, o + 1), nv, None)]
        if g + 1 < G:
            return [((t, "slot", g + 1, 0), nv, None)]
        return [((t + 1, "done", 0, 0) if t + 1 > T else (t + 1, "wait", 0, 0), nv, None)]

    def owner_moves(oi, loc, vals):
        t, phase, g, i, mask = loc
        if phase == "done":
            return None
        if phase == "waitpush":
            ready = [w for w in range(W) if get(vals, ("push", oi, g, w)) >= t]
            ahead = [w for w in ready if get(vals, ("push", oi, g, w)) > t]
            if ahead:
                return [(loc, vals, "a worker is ahead of an owner")]
            if len(ready) < need:
                return []
            return [((t, "check", g, 0, m), vals, None) for m in combinations(ready, need)]
        if phase == "check":
            if get(vals, ("slot", oi, g, mask[i])) != t:
                return [(loc, vals, "owner consumed a slot of another step")]
            if i + 1 < need:
                return [((t, "check", g, i + 1, mask), vals, None)]
            return [((t, "waitpush", g + 1, 0, ()), vals, None)] if g + 1 < G else [((t, "publish", 0, 0, ()), vals, None)]
        nv = put(vals, ("pf", i, oi), t + 1)                       # phase == "publish", i = rank
        if i + 1 < W:
            return [((t, "publish", 0, i + 1, ()), nv, None)]
        return [((t + 1, "done", 0, 0, ()) if t + 1 > T else (t + 1, "waitpush", 0, 0, ()), nv, None)]

    start = (tuple([(1, "wait", 0, 0)] * W + [(1, "waitpush", 0, 0, ())] * nO), tuple(mem0[k] for k in keys))
    seen, todo = {start}, deque([start])
    while todo:
        locs, vals = todo.popleft()
        moved, running = False, False
        for p, loc in enumerate(locs):
            moves = worker_moves(p, loc, vals) if p < W else owner_moves(p - W, loc, vals)
            if moves is None:
                continue
            running = True
            for nloc, nvals, err in moves:
                if err:
                    return {"error": err, "states": len(seen)}
                moved = True
                nxt = (locs[:p] + (nloc,) + locs[p + 1:], nvals)
                if nxt not in seen:
                    seen.add(nxt)
                    todo.append(nxt)
                    if len(seen) > limit:
                        return {"error": "state space larger than %d" % limit, "states": len(seen)}
        if running and not moved:
            return {"error": "deadlock", "states": len(seen), "at": locs}
    return {"error": None, "states": len(seen)}


@pytest.mark.parametrize("W,G,owners,T,need", [
    (3, 1, [0], 3, 2),        # centralized PS, 2 of 3
    (3, 2, [0], 2, 2),        # two backward groups: the mask may differ per group
    (2, 1, [0, 1], 3, 1),     # sharded PS, 1 of 2: owners themselves can be the stragglers
    (3, 1, [0, 1], 2, 2),     # sharded over two owners, 2 of 3
    (3, 2, [0, 1], 2, 2),     # ... with two backward groups (118 k states)
    (4, 1, [0], 2, 3),        # 3 of 4
])
def test_backup_worker_protocol_every_interleaving(W, G, owners, T, need):
    res = explore_backup(W, G, owners, T, need)
    assert res["error"] is None, res
    assert res["states"] > 100


def test_skip_ahead_is_an_optimisation_not_a_safety_requirement():
    """A straggler that replays the steps it missed (no skip-ahead) pushes flags that are too old to be used and
    eventually catches up or finishes: still no deadlock and no slot of another step consumed.  Skipping ahead only
    saves it the wasted work."""
    res = explore_backup(3, 1, [0], 3, 2, bug="no_skip_ahead")
    assert res["error"] is None, res
