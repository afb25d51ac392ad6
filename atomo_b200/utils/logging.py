"""Log-line formats.  The worker line is a de-facto API: the LR-tuning parser
greps/regex-parses it (``/root/reference/src/tiny_tuning_parser.py:13-27``,
format at ``distributed_worker.py:255-258``); the PS line is
``sync_replicas_master_nn.py:221``."""

WORKER_FMT = ("Worker: {}, Step: {}, Epoch: {} [{}/{} ({:.0f}%)], Loss: {:.4f}, Time Cost: {:.4f}, "
              "Comp: {:.4f}, Encode: {: .4f}, Comm: {: .4f}, Msg(MB): {: .4f}, Prec@1: {: .4f}, Prec@5: {: .4f}")
MASTER_FMT = "Master: Step: {}, Decode Cost: {}, Cur lr {}, Gather: {}"
TEST_FMT = "Test set: Step: {}, Average loss: {:.4f}, Prec@1: {} Prec@5: {}"


def worker_line(rank, step, epoch, seen, total, loss, time_cost, comp, encode, comm, msg_mb, prec1, prec5):
    pct = 100.0 * seen / max(total, 1)
    return WORKER_FMT.format(rank, step, epoch, seen, total, pct, loss, time_cost, comp, encode, comm,
                             msg_mb, prec1, prec5)


def master_line(step, decode_cost, lr, gather):
    return MASTER_FMT.format(step, decode_cost, lr, gather)


def test_line(step, loss, prec1, prec5):
    return TEST_FMT.format(step, loss, prec1, prec5)
