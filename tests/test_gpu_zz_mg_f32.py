"""GPU tests of GF_OPT_MG_MATRIX_PRECISION = 1: the multigrid V-cycle streams FP32 copies of the
level operators (half the HBM bytes per smoother / residual application).

The V-cycle replaces the reference's SSOR preconditioner (nonlinear_elasticity.cc:1180-1182,
linear_elasticity.cc:548-549). The outer CG still applies the FP64 matrix and stops on the FP64
residual (nonlinear:1171-1187, linear:540-552), so Newton counts and displacements must not move;
only CG iteration counts may differ by an iteration. Vectors and accumulation stay FP64.
"""
import numpy as np
import pytest

from helpers import lin_params, nl_params, rel_err
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi, multigrid, solvers
    build.build_cuda()
    capi.lib()
    return capi, solvers, multigrid


def assembled(libs, dim, reps, degree, numbering="lexicographic"):
    capi, solvers, mg = libs
    p = nl_params(poly_degree=degree, type_lin="CG")
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    H = mg.Hierarchy(prob)
    h = H.fine
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
    rng = np.random.RandomState(3)
    u = 1e-3 * rng.uniform(-1, 1, prob.n_dofs)
    u[prob.constrained != 0] = 0
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.nl_begin_step()
    h.nl_newton_assemble()
    return prob, H


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (3, 2, [4, 8, 4], "lexicographic"),
    (3, 1, [6, 10, 4], "cellwise"),
    (2, 2, [8, 16], "component_wise"),
    (2, 1, [10, 6], "lexicographic"),
])
def test_f32_operator_copy_matches_fp64_to_single_precision(libs, dim, degree, reps, numbering):
    capi, solvers, mg = libs
    prob, H = assembled(libs, dim, reps, degree, numbering)
    h = H.fine
    rng = np.random.RandomState(11)
    x = rng.uniform(-1, 1, prob.n_dofs)
    h.set_vector(capi.VEC_SCRATCH0, x)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y64 = h.get_vector(capi.VEC_SCRATCH1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y32 = h.get_vector(capi.VEC_SCRATCH1)
    # entries rounded to FP32 (2^-24 relative each), FP64 accumulation
    rowptr, col, val = h.export_csr(capi.MAT_TANGENT)
    import scipy.sparse as sp
    absA = sp.csr_matrix((np.abs(val), col, rowptr), shape=(prob.n_dofs, prob.n_dofs))
    bound = 2.0 ** -24 * (absA @ np.abs(x)) + 1e-13 * np.abs(y64).max()
    assert np.all(np.abs(y32 - y64) <= bound)
    assert rel_err(y32, y64) > 0.0          # it really is the single-precision copy
    # TMA-tiled and LDG kernels stream the same copy in the same order
    h.set_option(capi.OPT_SPMV_KERNEL, 1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    assert np.array_equal(h.get_vector(capi.VEC_SCRATCH1), y32)
    h.set_option(capi.OPT_SPMV_KERNEL, 0)
    # switching the option off releases the copy
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 0)
    with pytest.raises(capi.GraftError):
        h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    H.close()


def test_f32_vcycle_is_symmetric_positive_definite_and_close_to_fp64(libs):
    capi, solvers, mg = libs
    prob, H = assembled(libs, 3, [4, 8, 4], 2)
    h = H.fine
    rng = np.random.RandomState(5)
    free = prob.constrained == 0
    b1 = rng.uniform(-1, 1, prob.n_dofs) * free
    b2 = rng.uniform(-1, 1, prob.n_dofs) * free

    def vcycle(b):
        h.set_vector(capi.VEC_SCRATCH0, b)
        h.mg_vcycle(capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
        return h.get_vector(capi.VEC_SCRATCH1)

    z1, z2 = vcycle(b1), vcycle(b2)
    assert abs(b1 @ z2 - b2 @ z1) <= 1e-8 * max(abs(b1 @ z2), abs(b2 @ z1))
    assert b1 @ z1 > 0 and b2 @ z2 > 0
    assert rel_err(vcycle(b1 + 2 * b2), z1 + 2 * z2) < 1e-10      # linear in FP64 vectors
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 0)
    z1_64 = vcycle(b1)
    assert 0.0 < rel_err(z1, z1_64) < 1e-2
    H.close()


def test_f32_vcycle_keeps_newton_counts_and_displacements(libs):
    capi, solvers, mg = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                  max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    out = {}
    for prec in (0, 1):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_MG_MATRIX_PRECISION, prec)
        part = solvers.FakeParticipant(3, 3, p.delta_t, traction, 2)   # implicit, k = 2
        solid = solvers.Solid(prob, part, handle=H.fine)
        solid.run()
        out[prec] = ([[r[0] for r in rows] for rows in solid.history],
                     [d for (w, it, d) in part.written])
        H.close()
    its64, its32 = out[0][0], out[1][0]
    assert [len(r) for r in its32] == [len(r) for r in its64]         # identical Newton counts
    for a, b in zip(its32, its64):
        assert all(abs(x - y) <= 1 for x, y in zip(a, b))             # CG iterations +-1
    for d32, d64 in zip(out[1][1], out[0][1]):
        assert rel_err(d32, d64) < 1e-7


def test_f32_vcycle_linear_model(libs):
    capi, solvers, mg = libs
    p = lin_params(poly_degree=1, type_lin="CG", max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4])
    n = prob.n_iface_nodes
    buf = np.tile([300.0, -100.0, 50.0], n)
    out = {}
    for prec in (0, 1):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_MG_MATRIX_PRECISION, prec)
        part = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: buf)
        ed = solvers.ElastoDynamics(prob, part, handle=H.fine)
        ed.run()
        out[prec] = (ed.history, part.written[-1][2])
        H.close()
    assert all(res <= 1e-10 for it, res in out[1][0])
    assert all(abs(a[0] - b[0]) <= 1 for a, b in zip(out[1][0], out[0][0]))
    assert np.abs(out[1][1] - out[0][1]).max() < 1e-9


# ---- GF_OPT_MG_MATRIX_PRECISION = 2: x staged and accumulated in FP32 as well -----------------


@pytest.mark.parametrize("dim,degree,reps", [(3, 2, [4, 8, 4]), (3, 1, [6, 10, 4]), (2, 2, [8, 16])])
def test_all_fp32_operator_is_a_single_precision_spmv(libs, dim, degree, reps):
    capi, solvers, mg = libs
    prob, H = assembled(libs, dim, reps, degree)
    h = H.fine
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 2)
    x = np.random.RandomState(11).uniform(-1, 1, prob.n_dofs)
    h.set_vector(capi.VEC_SCRATCH0, x)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y64 = h.get_vector(capi.VEC_SCRATCH1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y32 = h.get_vector(capi.VEC_SCRATCH1)
    rowptr, col, val = h.export_csr(capi.MAT_TANGENT)
    import scipy.sparse as sp
    absA = sp.csr_matrix((np.abs(val), col, rowptr), shape=(prob.n_dofs, prob.n_dofs))
    n_max = np.diff(rowptr).max()
    # rounding of A, of x, and of every partial sum (2 chains + 5 shuffle levels per lane)
    bound = 2.0 ** -24 * (n_max / 32 + 12) * (absA @ np.abs(x)) + 1e-12 * np.abs(y64).max()
    assert np.all(np.abs(y32 - y64) <= bound)
    assert y32.astype(np.float32).astype(np.float64).tolist() == y32.tolist()   # FP32 results
    H.close()


def test_all_fp32_vcycle_keeps_newton_counts_and_displacements(libs):
    capi, solvers, mg = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                  max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    out = {}
    for prec in (0, 2):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_MG_MATRIX_PRECISION, prec)
        part = solvers.FakeParticipant(3, 3, p.delta_t, traction, 2)
        solid = solvers.Solid(prob, part, handle=H.fine)
        solid.run()
        out[prec] = ([[r[0] for r in rows] for rows in solid.history],
                     [d for (w, it, d) in part.written])
        H.close()
    assert [len(r) for r in out[2][0]] == [len(r) for r in out[0][0]]
    for a, b in zip(out[2][0], out[0][0]):
        assert all(abs(x - y) <= 2 for x, y in zip(a, b))
    for d32, d64 in zip(out[2][1], out[0][1]):
        assert rel_err(d32, d64) < 1e-7
