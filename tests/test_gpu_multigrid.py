"""GPU tests of the geometric multigrid preconditioner (gf_mg_attach / GF_PRECOND_MULTIGRID).

The reference preconditions CG with SSOR (nonlinear_elasticity.cc:1180-1182, linear_elasticity.cc:
548-549); any SPD preconditioner yields the same converged solution, so the checks are: the V-cycle
is symmetric positive definite, the preconditioned CG reproduces the oracle's Newton counts and
interface displacements, and the iteration count is far below block-Jacobi and mesh independent.
"""
import numpy as np
import pytest

from helpers import lin_params, nl_params, rel_err
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi, multigrid, solvers
    from oracle import oracle_py as orc
    build.build_cuda()
    capi.lib()
    return capi, solvers, multigrid, orc


def nl_assembled_hierarchy(libs, dim, reps, degree=2, n_levels=None, numbering="lexicographic"):
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=degree, type_lin="CG")
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    H = mg.Hierarchy(prob, n_levels=n_levels)
    h = H.fine
    rng = np.random.RandomState(3)
    u = 1e-3 * rng.uniform(-1, 1, prob.n_dofs)
    u[prob.constrained != 0] = 0
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.nl_begin_step()
    h.nl_newton_assemble()
    return prob, H


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (3, 2, [4, 8, 4], "lexicographic"),
    (3, 1, [4, 8, 4], "cellwise"),
    (2, 2, [8, 16], "component_wise"),
])
def test_vcycle_is_symmetric_positive_definite(libs, dim, degree, reps, numbering):
    capi, solvers, mg, orc = libs
    prob, H = nl_assembled_hierarchy(libs, dim, reps, degree, numbering=numbering)
    h = H.fine
    assert H.n_levels >= 2
    rng = np.random.RandomState(5)
    free = prob.constrained == 0
    vecs = []
    for k in range(2):
        b = rng.uniform(-1, 1, prob.n_dofs) * free
        h.set_vector(capi.VEC_SCRATCH0, b)
        h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
        vecs.append((b, h.get_vector(capi.VEC_SCRATCH1)))
    (b1, z1), (b2, z2) = vecs
    assert abs(b1 @ z2 - b2 @ z1) <= 1e-10 * max(abs(b1 @ z2), abs(b2 @ z1))
    assert b1 @ z1 > 0 and b2 @ z2 > 0
    assert np.all(z1[~free] == 0.0)
    # the V-cycle is linear: M(b1 + 2 b2) = M b1 + 2 M b2
    h.set_vector(capi.VEC_SCRATCH0, b1 + 2 * b2)
    h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    assert rel_err(h.get_vector(capi.VEC_SCRATCH1), z1 + 2 * z2) < 1e-11
    H.close()


def test_vcycle_approximates_the_inverse(libs):
    """||I - M A|| < 1 in the energy norm: one V-cycle reduces the error of A x = b markedly."""
    capi, solvers, mg, orc = libs
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    prob, H = nl_assembled_hierarchy(libs, 3, [4, 8, 4])
    h = H.fine
    rowptr, col, val = h.export_csr(capi.MAT_TANGENT)
    A = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs))
    rng = np.random.RandomState(7)
    x_true = rng.uniform(-1, 1, prob.n_dofs) * (prob.constrained == 0)
    b = A @ x_true
    h.set_vector(capi.VEC_SCRATCH0, b)
    h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    e = x_true - h.get_vector(capi.VEC_SCRATCH1)
    assert np.sqrt(e @ (A @ e)) < 0.5 * np.sqrt(x_true @ (A @ x_true))
    H.close()


def test_multigrid_cg_newton_counts_and_displacement_match_oracle(libs):
    """Coupled run, 'Solver type = CG': multigrid-CG on the device vs the oracle's SSOR-CG."""
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01)
    prob = make_problem(p, 3, reps=[4, 8, 2])
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1200.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    H = mg.Hierarchy(prob)
    assert H.n_levels == 2
    part = solvers.FakeParticipant(3, 3, p.delta_t, traction)
    solid = solvers.Solid(prob, part, handle=H.fine)
    solid.run()
    o = orc.Oracle(prob)
    counts, written = [], []
    for w in range(3):
        o.format_precice_to_deal(traction((w + 1) * p.delta_t, 0), orc.NL_EXTERNAL_STRESS)
        k, hist = o.nl_timestep()
        counts.append(k)
        written.append(o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT))
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-6   # both stop at 1e-6 relative linear residual
    H.close()


def test_multigrid_cg_iterations_are_mesh_independent_and_few(libs):
    capi, solvers, mg, orc = libs
    its = {}
    for reps in ([4, 16, 4], [8, 32, 8]):
        p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                      max_iterations_lin=1.0)
        prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
        n = prob.n_iface_nodes
        buf = np.tile([2000.0, 0.0, 0.0], n)
        out = {}
        for kind in ("jacobi", "mg"):
            H = mg.Hierarchy(prob) if kind == "mg" else None
            h = H.fine if H else capi.Handle(prob)
            part = solvers.FakeParticipant(3, 1, p.delta_t, lambda t, it: buf)
            solid = solvers.Solid(prob, part, handle=h)
            solid.run()
            out[kind] = ([r[0] for r in solid.history[0]], part.written[-1][2])
            (H or h).close()
        assert len(out["mg"][0]) == len(out["jacobi"][0])          # same Newton count
        assert rel_err(out["mg"][1], out["jacobi"][1]) < 1e-7
        assert max(out["mg"][0]) * 8 < max(out["jacobi"][0])
        its[tuple(reps)] = max(out["mg"][0])
    assert its[(8, 32, 8)] <= its[(4, 16, 4)] + 6


def test_multigrid_linear_model_matches_block_jacobi(libs):
    capi, solvers, mg, orc = libs
    p = lin_params(poly_degree=1, type_lin="CG", max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4])
    n = prob.n_iface_nodes
    buf = np.tile([300.0, -100.0, 50.0], n)
    out = {}
    for kind in ("jacobi", "mg"):
        H = mg.Hierarchy(prob) if kind == "mg" else None
        h = H.fine if H else capi.Handle(prob)
        part = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: buf)
        ed = solvers.ElastoDynamics(prob, part, handle=h)
        ed.run()
        out[kind] = (ed.history, part.written[-1][2])
        (H or h).close()
    assert all(res <= 1e-10 for it, res in out["mg"][0])
    assert max(it for it, res in out["mg"][0]) * 3 < max(it for it, res in out["jacobi"][0])
    assert np.abs(out["mg"][1] - out["jacobi"][1]).max() < 1e-9


def test_attach_rejects_bad_hierarchies(libs):
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=2)
    fine = make_problem(p, 3, reps=[4, 4, 4], numbering="lexicographic")
    coarse = mg.coarsen_problem(fine)
    hf, hc = capi.Handle(fine), capi.Handle(coarse)
    tab = mg.child_table(coarse.mesh, fine.mesh)
    bad = tab.copy()
    bad[0, 0] = bad[0, 1]                      # a fine cell with two parents
    with pytest.raises(capi.GraftError):
        hf.mg_attach(hc, bad)
    hf.set_option(capi.OPT_PRECONDITIONER, capi.PRECOND_MULTIGRID)
    hf.nl_begin_step()
    hf.nl_newton_assemble()
    with pytest.raises(capi.GraftError):       # multigrid selected, nothing attached
        hf.nl_newton_solve(0, 1e-6, 1.0)
    hf.mg_attach(hc, tab)
    hf.nl_newton_assemble()
    hf.nl_newton_solve(0, 1e-6, 1.0)
    # levels may be destroyed in any order
    hc.close()
    hf.close()


def test_better_initial_guess_keeps_the_newton_history_and_saves_cg_iterations(libs):
    """GF_OPT_CG_INITIAL_GUESS: 0 = the reference's guess (previous newton_update,
    nonlinear_elasticity.cc:1184), 1 (default) = the zero vector whenever its residual is smaller.
    Same stopping criterion, so Newton counts and displacements agree within the solver tolerance;
    the default needs clearly fewer CG iterations from the second Newton pass on."""
    capi, solvers, mg = libs[:3]
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01, max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    out = {}
    for guess in (0, 1):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_CG_INITIAL_GUESS, guess)
        part = solvers.FakeParticipant(3, 2, p.delta_t, traction)
        solid = solvers.Solid(prob, part, handle=H.fine)
        solid.run()
        out[guess] = ([[int(r[0]) for r in rows] for rows in solid.history],
                      [d for (w, it, d) in part.written])
        H.close()
    assert [len(r) for r in out[0][0]] == [len(r) for r in out[1][0]]          # Newton counts
    for a, b in zip(out[0][1], out[1][1]):
        assert rel_err(a, b) < 1e-7
    for ref, new in zip(out[0][0], out[1][0]):
        assert new[0] == ref[0]                                                # first pass: same guess (zero)
        assert all(x <= y for x, y in zip(new, ref))
    assert sum(map(sum, out[1][0])) < 0.9 * sum(map(sum, out[0][0]))


def test_semi_coarsened_levels_below_a_large_coarsest_level(libs):
    """Repetitions 3 x 36 x 12 cannot be halved isotropically (3 is odd) and the level has 12,775
    nodes - the host hierarchy then halves the even directions only (3x18x6, 3x9x3) and
    gf_mg_attach reads the refined directions off the cell geometry. The V-cycle must stay a
    symmetric positive definite preconditioner, CG must need few iterations, and the Newton run
    must agree with the block-Jacobi CG on the same mesh."""
    capi, solvers, mg = libs[:3]
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01, max_iterations_lin=1.0,
                  tol_lin=1e-8)
    prob = make_problem(p, 3, reps=[3, 36, 12], numbering="lexicographic")
    H = mg.Hierarchy(prob)
    assert [q.mesh.reps for q in H.problems] == [[3, 36, 12], [3, 18, 6], [3, 9, 3]]
    n = prob.n_iface_nodes
    buf = np.tile([1500.0, 0.0, 50.0], n)
    out = {}
    for use_mg in (True, False):
        h = H.fine if use_mg else capi.Handle(prob)
        part = solvers.FakeParticipant(3, 1, p.delta_t, lambda t, it: buf)
        solid = solvers.Solid(prob, part, handle=h)
        solid.run()
        out[use_mg] = ([[int(r[0]) for r in rows] for rows in solid.history], part.written[0][2])
        if use_mg:
            # symmetry of the V-cycle: x . M y == y . M x
            rng = np.random.RandomState(1)
            free = prob.constrained == 0
            x, y = rng.uniform(-1, 1, prob.n_dofs) * free, rng.uniform(-1, 1, prob.n_dofs) * free
            h.set_vector(capi.VEC_SCRATCH0, x)
            h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
            Mx = h.get_vector(capi.VEC_SCRATCH1)
            h.set_vector(capi.VEC_SCRATCH0, y)
            h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
            My = h.get_vector(capi.VEC_SCRATCH1)
            assert abs(x @ My - y @ Mx) <= 1e-10 * abs(x @ Mx) and x @ Mx > 0
        else:
            h.close()
    H.close()
    assert [len(r) for r in out[True][0]] == [len(r) for r in out[False][0]]
    assert max(map(max, out[True][0])) <= 40 < min(map(min, out[False][0]))
    assert rel_err(out[True][1], out[False][1]) < 1e-7
