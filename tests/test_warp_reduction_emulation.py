"""CPU emulation of the two warp reductions of the SpMV consumers (csrc/kernel_utils.cuh):
warp_sum — the plain xor-butterfly, one per scalar row — and warp_sum_rows — ONE transposed
butterfly for the 2 or 3 scalar rows of a block row (6 instead of 15 shuffle rounds). The
transposed form combines the same partial sums in the same order (xor 16, 8, 4, 2, 1), only in
fewer lanes, so it must give BITWISE the same sums; shuffles are emulated as array indexing over
the 32 lanes, IEEE double addition by numpy."""
import numpy as np

LANES = np.arange(32)


def warp_sum(v):
    v = v.copy()
    for o in (16, 8, 4, 2, 1):
        v = v + v[LANES ^ o]               # v += __shfl_xor_sync(full, v, o)
    return v


def warp_sum_rows(vs):
    b4, b3 = (LANES & 16) != 0, (LANES & 8) != 0
    if len(vs) == 2:
        keep, send = np.where(b4, vs[1], vs[0]), np.where(b4, vs[0], vs[1])
        t = keep + send[LANES ^ 16]
        t = t + t[LANES ^ 8]
    else:
        zero = np.zeros(32)
        keep0, keep1 = np.where(b4, vs[2], vs[0]), np.where(b4, zero, vs[1])
        send0, send1 = np.where(b4, vs[0], vs[2]), np.where(b4, vs[1], zero)
        r0, r1 = keep0 + send0[LANES ^ 16], keep1 + send1[LANES ^ 16]
        t = np.where(b3, r1, r0) + np.where(b3, r0, r1)[LANES ^ 8]
    for o in (4, 2, 1):
        t = t + t[LANES ^ o]
    return t


def test_transposed_row_reduction_is_bitwise_the_butterfly():
    rng = np.random.RandomState(0)
    for trial in range(3000):
        dim = 2 + trial % 2
        vs = [rng.uniform(-1, 1, 32) * 10.0 ** rng.randint(-12, 12, 32) for _ in range(dim)]
        t = warp_sum_rows(vs)
        stride = 8 if dim == 3 else 16
        for r in range(dim):
            ref = warp_sum(vs[r])
            assert np.all(ref == ref[0])                       # every lane ends with the same bits
            assert np.all(t[r * stride:(r + 1) * stride] == ref[0])


def test_fused_dot_partials_keep_their_summation_order():
    """The lanes holding the row sums move from 0,1,2 to 0,8,16 (0,16): the final warp_sum over the
    per-lane dot partials must still group them as ((d0 + d2) + d1) resp. (d0 + d1)."""
    rng = np.random.RandomState(1)
    for _ in range(2000):
        d = rng.uniform(-1, 1, 3) * 10.0 ** rng.randint(-10, 10, 3)
        plain, moved = np.zeros(32), np.zeros(32)
        plain[:3] = d
        moved[[0, 8, 16]] = d
        assert warp_sum(plain)[0] == warp_sum(moved)[0]
        plain2, moved2 = np.zeros(32), np.zeros(32)
        plain2[:2] = d[:2]
        moved2[[0, 16]] = d[:2]
        assert warp_sum(plain2)[0] == warp_sum(moved2)[0]
