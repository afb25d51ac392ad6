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
