"""Deterministic unit tests of the backup-round state machine (parallel/backup_rounds.py): events are fed through
the queue by hand, sends are recorded instead of performed.  The multi-process behaviour (real gloo, a real straggler,
real process deaths) is covered by tests/test_runtime_cpu.py."""
import queue

import pytest

from atomo_b200.parallel.backup_rounds import BackupRounds


class _T:
    world_size = 4
    device = "cpu"
    group = None
    _stale_dropped = 0


def make(need, workers=(1, 2, 3)):
    b = BackupRounds.__new__(BackupRounds)
    b.t, b.need, b.workers = _T(), need, list(workers)
    b.sent = {w: 0 for w in workers}
    b.recvd = {w: 0 for w in workers}
    b.dead, b.byes, b.asked = set(), set(), set()
    b.inflight = {w: [] for w in workers}
    b.round, b.q, b._rx_expect, b._threads = None, queue.Queue(), None, []
    b.announced = []

    def announce(w):
        b.sent[w] += 1
        b.asked.add(w)
        b.announced.append((b.round[0], w))
    b._announce = announce
    return b


def start_round(b, step):
    import torch
    return b.send_round(step, torch.zeros(4))


def msg(b, w, step, codes="g"):
    b.q.put(("msg", w, (step, {"codes": codes})))


UNPACK = lambda buf: buf


def test_round_goes_to_free_workers_and_gather_returns_after_need():
    b = make(need=2)
    assert start_round(b, 1) == [1, 2, 3]
    msg(b, 3, 1); msg(b, 1, 1)
    got = b.gather(1, UNPACK)
    assert sorted(got) == [1, 3] and b.recvd == {1: 1, 2: 0, 3: 1}
    # worker 2 still owes step 1: the next round goes to 1 and 3 only
    assert start_round(b, 2) == [1, 3]


def test_late_message_is_dropped_and_its_sender_joins_the_current_step():
    b = make(need=2)
    start_round(b, 1)
    msg(b, 1, 1); msg(b, 3, 1)
    b.gather(1, UNPACK)
    start_round(b, 2)
    msg(b, 2, 1)            # the straggler's step-1 gradient surfaces during step 2
    msg(b, 2, 2); msg(b, 1, 2)
    got = b.gather(2, UNPACK)
    assert (2, 2) in b.announced             # announced step 2 the moment it became free
    assert sorted(got) == [1, 2] and b.t._stale_dropped == 1


def test_send_round_waits_until_need_workers_are_free():
    b = make(need=2)
    start_round(b, 1)
    msg(b, 1, 1); msg(b, 2, 1)
    b.gather(1, UNPACK)
    start_round(b, 2)                         # to 1 and 2; 3 still owes step 1
    msg(b, 1, 2); msg(b, 3, 1); msg(b, 3, 2)
    got = b.gather(2, UNPACK)                 # 1 and late-joiner 3 deliver; 2 now owes step 2
    assert sorted(got) == [1, 3]
    assert start_round(b, 3) == [1, 3]        # two free workers: no waiting
    b2 = make(need=3)
    start_round(b2, 1)
    for w in (1, 2, 3):
        msg(b2, w, 1)
    b2.gather(1, UNPACK)
    start_round(b2, 2)
    msg(b2, 1, 2); msg(b2, 2, 2); msg(b2, 3, 2)
    assert sorted(b2.gather(2, UNPACK)) == [1, 2, 3]


def test_lost_worker_shrinks_the_requirement(capsys):
    b = make(need=2)
    start_round(b, 1)
    msg(b, 1, 1)
    b.q.put(("lost", 2, "Connection closed by peer"))
    b.q.put(("lost", 3, "Connection closed by peer"))
    got = b.gather(1, UNPACK)                 # 2 and 3 are gone: one gradient is all there will ever be
    assert sorted(got) == [1] and b.dead == {2, 3}
    assert "worker 2 is gone" in capsys.readouterr().out
    assert start_round(b, 2) == [1]           # need is capped by the number of survivors


def test_everyone_asked_is_gone_but_a_straggler_saves_the_step():
    b = make(need=1)
    start_round(b, 1)
    msg(b, 1, 1)
    b.gather(1, UNPACK)                       # 2 and 3 still owe step 1
    start_round(b, 2)                         # only worker 1 is asked
    b.q.put(("lost", 1, "closed"))
    msg(b, 3, 1)                              # straggler frees up -> late joiner of step 2
    msg(b, 3, 2)
    got = b.gather(2, UNPACK)
    assert sorted(got) == [3]


def test_all_workers_gone_is_an_error():
    b = make(need=1, workers=(1,))
    start_round(b, 1)
    b.q.put(("lost", 1, "closed"))
    with pytest.raises(RuntimeError, match="every worker is gone"):
        b.gather(1, UNPACK)
