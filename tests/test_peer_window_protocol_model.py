"""Model of the peer-window transport (csrc/comm.cu) on the CPU: every rank is a coroutine that runs
its stream of operations in order — halo push (remote stores into the neighbour's mailbox, then a
release store of the pair epoch into the neighbour's flag), halo wait (acquire-spin on the own flag
until it reaches the epoch, then unpack) and the single-kernel all-reduce (scatter my values into
every rank's slot, release the flags, acquire all flags in my window, add in rank order) — with
mailboxes, slots and flags double-buffered by the parity of the epoch exactly as the kernels
address them. Ranks are interleaved by a random scheduler (a rank may run far ahead of another).

Checked: no rank ever unpacks a mailbox or sums a slot that does not hold the data of exactly the
awaited epoch (no overwrite before the reader is done, no stale read), no deadlock, and the
all-reduce result is bitwise the same on every rank. A single-buffered variant is shown to fail, so
the model really exercises the hazard the double buffering removes."""
import random

import pytest


class Window:
    """One rank's window: [parity][sender] mailboxes / flags, all-reduce slots / flags."""

    def __init__(self, n_ranks, n_buffers):
        self.nb = n_buffers
        self.mbox = [[None] * n_ranks for _ in range(n_buffers)]
        self.hflag = [[0] * n_ranks for _ in range(n_buffers)]
        self.slot = [[None] * n_ranks for _ in range(n_buffers)]
        self.aflag = [[0] * n_ranks for _ in range(n_buffers)]


def run_ranks(n_ranks, n_steps, seed, n_buffers=2, skew=4):
    rng = random.Random(seed)
    win = [Window(n_ranks, n_buffers) for _ in range(n_ranks)]
    nbrs = {r: [x for x in (r - 1, r + 1) if 0 <= x < n_ranks] for r in range(n_ranks)}   # slabs
    results = [[] for _ in range(n_ranks)]
    # the same random operation sequence on every rank (collective calls)
    ops = [rng.choice(("halo", "reverse_halo", "allreduce")) for _ in range(n_steps)]

    def rank_program(me):
        halo_epoch = {r: 0 for r in nbrs[me]}
        ar_epoch = 0
        for step, op in enumerate(ops):
            if op in ("halo", "reverse_halo"):
                epochs = {}
                for r in nbrs[me]:                                   # halo_push_kernel
                    halo_epoch[r] += 1
                    e = epochs[r] = halo_epoch[r]
                    par = e % n_buffers
                    win[r].mbox[par][me] = ("data", me, r, e)        # remote stores
                    yield
                    win[r].hflag[par][me] = e                        # st.release.sys
                    yield
                for r in nbrs[me]:                                   # halo_wait_kernel
                    e = epochs[r]
                    par = e % n_buffers
                    while win[me].hflag[par][r] < e:                 # ld.acquire.sys spin
                        yield
                    got = win[me].mbox[par][r]
                    assert got == ("data", r, me, e), \
                        "rank %d unpacked %r, expected epoch %d from rank %d" % (me, got, e, r)
                    yield
                    assert win[me].mbox[par][r] == ("data", r, me, e), "mailbox overwritten while unpacking"
            else:                                                    # p2p_allreduce_kernel
                ar_epoch += 1
                par = ar_epoch % n_buffers
                mine = (me + 1) * 0.1 * ar_epoch
                for r in range(n_ranks):
                    win[r].slot[par][me] = (ar_epoch, mine)
                    yield
                for r in range(n_ranks):
                    win[r].aflag[par][me] = ar_epoch
                yield
                for r in range(n_ranks):
                    while win[me].aflag[par][r] < ar_epoch:
                        yield
                s = 0.0
                for r in range(n_ranks):                             # rank order: same bits everywhere
                    e, v = win[me].slot[par][r]
                    assert e == ar_epoch, "rank %d summed the slot of epoch %d in epoch %d" % (me, e, ar_epoch)
                    s += v
                    yield
                results[me].append(s)
            yield

    progs = [rank_program(r) for r in range(n_ranks)]
    alive = list(range(n_ranks))
    idle = 0
    while alive:
        r = rng.choice(alive)
        for _ in range(rng.randint(1, skew * 8)):                    # let one rank run ahead
            try:
                next(progs[r])
            except StopIteration:
                alive.remove(r)
                break
        idle += 1
        assert idle < 2_000_000, "deadlock"
    assert all(res == results[0] for res in results)
    return results[0]


@pytest.mark.parametrize("n_ranks", [2, 3, 4, 8])
def test_double_buffered_peer_windows_never_mix_epochs(n_ranks):
    for seed in range(40):
        run_ranks(n_ranks, n_steps=60, seed=seed)
        run_ranks(n_ranks, n_steps=60, seed=1000 + seed, skew=40)    # very uneven progress


def test_single_buffering_would_be_overwritten():
    failures = 0
    for seed in range(60):
        try:
            run_ranks(3, n_steps=60, seed=seed, n_buffers=1, skew=40)
        except AssertionError:
            failures += 1
    assert failures > 0
