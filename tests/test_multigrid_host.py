"""CPU tests of the host side of the multigrid hierarchy (dealii_adapter_b200/multigrid.py): level
problems, child tables (deal.II child order) and their restriction to slab partitions."""
import numpy as np

from helpers import nl_params
from dealii_adapter_b200 import multigrid as mg
from dealii_adapter_b200.problem import make_problem


def test_coarsen_keeps_geometry_and_boundary_roles(native_libs):
    p = nl_params(poly_degree=2)
    fine = make_problem(p, 3, reps=[4, 8, 2], numbering="lexicographic")
    coarse = mg.coarsen_problem(fine)
    assert coarse.mesh.reps == [2, 4, 1]
    assert coarse.mesh.p0 == fine.mesh.p0 and coarse.mesh.p1 == fine.mesh.p1
    assert mg.coarsen_problem(coarse) is None          # a repetition is odd
    # nested nodes: every coarse support point is a fine support point with the same constraint
    fine_pts = {tuple(np.round(x, 12)): i for i, x in enumerate(fine.mesh.support_points[::3])}
    fz = fine.constrained.reshape(-1, 3)
    for i, x in enumerate(coarse.mesh.support_points[::3]):
        j = fine_pts[tuple(np.round(x, 12))]
        assert np.array_equal(fz[j], coarse.constrained.reshape(-1, 3)[i])
    cz = coarse.constrained.reshape(-1, 3)
    pts = coarse.mesh.support_points[::3]
    assert np.all(cz[np.isclose(pts[:, 1], 0.0)] == 1)            # clamped bottom, all components


def test_child_table_is_deal_ii_child_order(native_libs):
    p = nl_params(poly_degree=1)
    for dim, reps in ((2, [4, 6]), (3, [2, 4, 6])):
        fine = make_problem(p, dim, reps=reps)
        coarse = mg.coarsen_problem(fine)
        tab = mg.child_table(coarse.mesh, fine.mesh)
        assert sorted(tab.reshape(-1)) == list(range(fine.mesh.n_cells))
        nv = 1 << dim
        fv = fine.mesh.cell_vertices.reshape(-1, nv, dim)
        cv = coarse.mesh.cell_vertices.reshape(-1, nv, dim)
        for pc in range(coarse.mesh.n_cells):
            lo, hi = cv[pc, 0], cv[pc, nv - 1]
            for k in range(nv):
                off = np.array([(k >> d) & 1 for d in range(dim)])
                assert np.allclose(fv[tab[pc, k], 0], lo + 0.5 * off * (hi - lo))


def test_partitioned_child_tables_cover_the_local_fine_cells(native_libs):
    p = nl_params(poly_degree=2)
    fine = make_problem(p, 3, reps=[2, 16, 2], numbering="lexicographic")
    coarse = mg.coarsen_problem(fine)
    tab = mg.child_table(coarse.mesh, fine.mesh)
    world = 2
    for rank in range(world):
        pf, pc = fine.mesh.partition(1, world, rank), coarse.mesh.partition(1, world, rank)
        g2l = -np.ones(fine.mesh.n_cells, dtype=np.int64)
        g2l[pf.local_cell_global] = np.arange(pf.n_local_cells)
        local = g2l[tab[pc.local_cell_global]]
        covered = np.zeros(pf.n_local_cells, dtype=bool)
        covered[local[local >= 0]] = True
        assert covered.all()
        # the coarse ghost layer spans two fine layers, only the first is local on the fine level
        if rank == 0:
            assert (local < 0).any()


def test_level_plan_for_partitioned_runs_replicates_the_small_levels(native_libs):
    """world > 1: levels stay slab-partitioned while the slabs of consecutive levels coincide and the
    level is large; the small levels are replicated on every rank (no coarse halo exchanges)."""
    from dealii_adapter_b200 import multigrid as mg
    p = nl_params(poly_degree=2, scenario="PF")
    prob = make_problem(p, 3, reps=[8, 32, 8])
    # serial: nothing is replicated, coarsening stops at the first odd repetition
    probs, rep = mg.plan_levels(prob)
    assert [q.mesh.reps for q in probs] == [[8, 32, 8], [4, 16, 4], [2, 8, 2], [1, 4, 1]]
    assert rep == [False] * 4
    # 2 ranks, everything below 20k dofs replicated
    probs, rep = mg.plan_levels(prob, world=2, axis=1, replicate_below_dofs=20000)
    assert [q.n_dofs for q in probs][1] == 9 * 33 * 9 * 3 and rep == [False, True, True, True]
    # 4 ranks, no size threshold: 32 -> 16 -> 8 -> 4 layers = 8, 4, 2, 1 per rank: all aligned
    probs, rep = mg.plan_levels(prob, world=4, axis=1, replicate_below_dofs=0)
    assert rep == [False, False, False, False]
    # 8 ranks: 4 -> 2 -> 1 layers per rank are aligned, half a layer per rank is not
    probs, rep = mg.plan_levels(prob, world=8, axis=1, replicate_below_dofs=0)
    assert rep == [False, False, False, True]
    # 3 ranks: 32 layers cannot be cut into aligned slabs at all -> every coarse level replicated
    probs, rep = mg.plan_levels(prob, world=3, axis=1, replicate_below_dofs=0)
    assert rep == [False, True, True, True]
    # once replicated, always replicated
    for w in (2, 3, 4, 8):
        _, rep = mg.plan_levels(prob, world=w, axis=1)
        assert rep == sorted(rep)


def test_semi_coarsening_below_a_large_level_with_an_odd_repetition(native_libs):
    """3 x 36 x 12 (12,775 Q2 nodes) cannot be halved in x: the even directions are halved alone
    until the level is small; children along the unrefined direction do not exist (-1)."""
    p = nl_params(poly_degree=2)
    fine = make_problem(p, 3, reps=[3, 36, 12], numbering="lexicographic")
    problems, replicated = mg.plan_levels(fine)
    assert [q.mesh.reps for q in problems] == [[3, 36, 12], [3, 18, 6], [3, 9, 3]]
    assert mg.coarsen_problem(problems[1], allow_semi=False) is None
    tab = mg.child_table(problems[1].mesh, fine.mesh)
    assert tab.shape == (3 * 18 * 6, 8)
    assert np.all(tab[:, 1::2] == -1)                               # kx = 1 children do not exist
    assert sorted(tab[:, 0::2].reshape(-1)) == list(range(fine.mesh.n_cells))
    fv = fine.mesh.cell_vertices.reshape(-1, 8, 3)
    cv = problems[1].mesh.cell_vertices.reshape(-1, 8, 3)
    for pc in (0, 17, 200):
        for k in (0, 2, 4, 6):
            lo, hi = fv[tab[pc, k], 0], fv[tab[pc, k], 7]
            assert np.all(lo >= cv[pc, 0] - 1e-12) and np.all(hi <= cv[pc, 7] + 1e-12)
            assert np.isclose(hi[0] - lo[0], cv[pc, 7][0] - cv[pc, 0][0])      # full width in x
    # small meshes keep the round-1 behaviour: no semi-coarsening below the threshold
    small = make_problem(p, 3, reps=[3, 18, 3], numbering="lexicographic")
    assert mg.coarsen_problem(small) is None
