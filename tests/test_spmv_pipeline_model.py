"""Model of the two-ring SpMV pipeline protocol (csrc/spmv.cu, spmv_tma2_kernel) run on the CPU
under random schedules: the producer warp, the gather groups and the consumer warps are
coroutines that block on mbarriers exactly where the kernel does; TMA copies complete after
random delays; mbarriers carry an arrival count, a transaction count and a phase, and a parity
wait passes iff the barrier's current phase parity differs from the awaited one (the hardware only
knows the parity, so a role that skips a phase of a barrier can alias).

Checked for many (value stages, x stages, gather groups, consumer warps, tiles per CTA): no
deadlock, every consumer reads tile k's values, column records and gathered x from stages that
hold exactly tile k at that moment, and the ordering rule of the producer (columns LEAD tiles
ahead of values) never throttles or deadlocks. The kernel's static_assert that the gather groups
must divide the x stages is shown to be necessary: the model catches the aliasing without it."""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        self._check()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def passed(self, parity):          # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Deadlock(Exception):
    pass


def simulate(n_my, vstages, xstages, groups, gather_warps, cons_warps, seed, max_delay=6):
    rng = random.Random(seed)
    lead = xstages - vstages
    vfull = [MBar(1) for _ in range(vstages)]
    vempty = [MBar(cons_warps) for _ in range(vstages)]
    cfull = [MBar(1) for _ in range(xstages)]
    xfull = [MBar(gather_warps // groups) for _ in range(xstages)]
    xempty = [MBar(cons_warps) for _ in range(xstages)]
    vslot = [None] * vstages           # which tile's values a stage holds (None while in flight)
    cslot = [None] * xstages           # column indices + row records
    xslot = [None] * xstages           # gathered x
    in_flight = []                     # (ready_time, callback)
    clock = [0]
    consumed = [[] for _ in range(cons_warps)]

    def tma(slot_list, s, tile, bar, nbytes):
        slot_list[s] = None            # being overwritten
        def done():
            slot_list[s] = tile
            bar.complete_tx(nbytes)
        in_flight.append((clock[0] + rng.randint(1, max_delay), done))

    def wait(bar, parity):
        while not bar.passed(parity):
            yield

    def producer():
        for i in range(n_my + lead):
            if i < n_my:
                sx = i % xstages
                if i >= xstages:
                    yield from wait(xempty[sx], ((i // xstages) - 1) & 1)
                cfull[sx].arrive_expect_tx(2)
                xslot[sx] = None       # the stage is recycled: its gathered x is stale from now on
                tma(cslot, sx, i, cfull[sx], 2)
            j = i - lead
            if j >= 0:
                sv = j % vstages
                if j >= vstages:
                    yield from wait(vempty[sv], ((j // vstages) - 1) & 1)
                vfull[sv].arrive_expect_tx(1)
                tma(vslot, sv, j, vfull[sv], 1)
            yield

    def gather(warp):
        grp = warp // (gather_warps // groups)
        for k in range(grp, n_my, groups):
            sx = k % xstages
            yield from wait(cfull[sx], (k // xstages) & 1)
            assert cslot[sx] == k, "gather warp read the columns of tile %r for tile %d" % (cslot[sx], k)
            yield                      # loads in flight
            xslot[sx] = k              # (every warp of the group writes its share)
            xfull[sx].arrive()
            yield

    def consumer(warp):
        for k in range(n_my):
            sv, sx = k % vstages, k % xstages
            yield from wait(cfull[sx], (k // xstages) & 1)
            yield from wait(xfull[sx], (k // xstages) & 1)
            yield from wait(vfull[sv], (k // vstages) & 1)
            assert cslot[sx] == k and xslot[sx] == k and vslot[sv] == k, \
                "consumer of tile %d saw records %r, x %r, values %r" % (k, cslot[sx], xslot[sx], vslot[sv])
            yield                      # FMAs
            assert cslot[sx] == k and xslot[sx] == k and vslot[sv] == k, "stage recycled under a consumer"
            consumed[warp].append(k)
            vempty[sv].arrive()
            xempty[sx].arrive()
            yield

    roles = [producer()] + [gather(w) for w in range(gather_warps)] + \
        [consumer(w) for w in range(cons_warps)]
    alive = list(range(len(roles)))
    idle_rounds = 0
    while alive:
        clock[0] += 1
        for item in [x for x in in_flight if x[0] <= clock[0]]:
            in_flight.remove(item)
            item[1]()
        state = (tuple(b.phase for b in vfull + vempty + cfull + xfull + xempty), len(in_flight))
        rng.shuffle(alive)
        for r in list(alive[:rng.randint(1, len(alive))]):
            try:
                next(roles[r])
            except StopIteration:
                alive.remove(r)
        new_state = (tuple(b.phase for b in vfull + vempty + cfull + xfull + xempty), len(in_flight))
        idle_rounds = idle_rounds + 1 if (new_state == state and not in_flight) else 0
        if idle_rounds > 2000:
            raise Deadlock("no progress: %d roles blocked" % len(alive))
    assert all(c == list(range(n_my)) for c in consumed)
    return clock[0]


KERNEL_CONFIGS = [          # (vstages, xstages, groups, gather warps, consumer warps) in spmv.cu
    (3, 4, 2, 8, 8),        # kind 2, FP64
    (3, 4, 2, 8, 16),       # kind 3 / 6, FP64 (default for plain launches)
    (3, 4, 1, 4, 16),       # kind 4, FP64
    (4, 6, 2, 8, 8),        # kind 2, FP32 copy
    (4, 6, 2, 8, 16),       # kind 3, FP32 copy and the all-FP32 operator
    (4, 6, 1, 4, 16),       # kind 4, FP32 copy
]


@pytest.mark.parametrize("cfg", KERNEL_CONFIGS)
def test_two_ring_protocol_is_deadlock_free_and_never_mixes_tiles(cfg):
    for n_my in (0, 1, 2, 3, 4, 5, 7, 8, 13, 24, 61):
        for seed in range(12):
            simulate(n_my, *cfg, seed=seed)
            simulate(n_my, *cfg, seed=1000 + seed, max_delay=40)    # slow TMA


def test_the_model_sees_the_parity_aliasing_of_groups_that_do_not_divide_the_stages():
    """4 gather groups on 6 x-stages (tried first, rejected): group 0 meets stage 0 at tiles 0, 12,
    24 - phases 0, 2, 4 of that barrier, skipping the odd ones - so its parity wait for tile 12 can
    pass on the completion of tile 0's phase while tile 6's columns are still in flight."""
    caught = 0
    for seed in range(300):
        try:
            simulate(40, 4, 6, 4, 8, 8, seed=seed, max_delay=60)
        except (AssertionError, Deadlock):
            caught += 1
    assert caught > 0
