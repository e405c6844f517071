"""GPU tests of the SpMV kernel kinds (GF_OPT_SPMV_KERNEL): the two-ring TMA kernels (3: 8 gather +
16 consumer warps; 6: the same with the transposed row reduction = the default 0) use the same
tiles and the same per-row summation order as the single-ring TMA kernel (5) and the LDG kernel
(1), so y = A x (vmult inside SolverCG, nonlinear_elasticity.cc:1184, linear_elasticity.cc:551)
must be BITWISE identical, for the FP64 matrix and for its FP32 V-cycle copy. The fused dot
product of the CG vmult is summed over the consumer warps of a CTA, so kinds with a different warp
count / lane layout may move the CG history in the last bits (never the iteration counts);
a kind always reproduces itself bit for bit."""
import os

import numpy as np
import pytest

from helpers import lin_params, nl_params, smooth_field
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi, multigrid, solvers
    build.build_cuda()
    capi.lib()
    return capi, solvers, multigrid


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (3, 2, [6, 12, 4], "lexicographic"),
    (3, 1, [10, 14, 6], "cellwise"),
    (2, 2, [16, 24], "component_wise"),
    (2, 1, [30, 18], "lexicographic"),
])
def test_two_ring_kernel_is_bitwise_equal(libs, dim, degree, reps, numbering):
    capi, solvers, mg = libs
    p = nl_params(poly_degree=degree, type_lin="CG")
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    H = mg.Hierarchy(prob)
    h = H.fine
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
    rng = np.random.RandomState(3)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, smooth_field(prob, 2e-3, seed=2))
    h.nl_begin_step()
    h.nl_newton_assemble()
    h.set_vector(capi.VEC_SCRATCH0, rng.uniform(-1, 1, prob.n_dofs))
    for mat in (capi.MAT_TANGENT, capi.MAT_MG_F32):
        ys = []
        kinds = (5, 0, 1, 3, 3, 6)            # twice: the rings are re-initialised per launch
        for kind in kinds:
            h.set_option(capi.OPT_SPMV_KERNEL, kind)
            h.spmv(mat, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
            ys.append(h.get_vector(capi.VEC_SCRATCH1))
        assert np.abs(ys[0]).max() > 0
        for y in ys[1:]:
            assert np.array_equal(y, ys[0])
    H.close()


def test_two_ring_kernel_reproduces_a_coupled_run_bit_for_bit(libs):
    capi, solvers, mg = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                  max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    out = {}
    for kind in (6, 0, 0):
        for prec in (0, 1):
            H = mg.Hierarchy(prob)
            H.fine.set_option(capi.OPT_SPMV_KERNEL, kind)
            H.fine.set_option(capi.OPT_MG_MATRIX_PRECISION, prec)
            part = solvers.FakeParticipant(3, 2, p.delta_t, traction, 2)
            solid = solvers.Solid(prob, part, handle=H.fine)
            solid.run()
            res = (np.array([r for rows in solid.history for r in rows]),
                   np.array([d for (w, it, d) in part.written]))
            H.close()
            if (kind, prec) in out:     # the default twice: run-to-run bitwise
                assert np.array_equal(res[0], out[kind, prec][0])
                assert np.array_equal(res[1], out[kind, prec][1])
            out[kind, prec] = res
    for prec in (0, 1):                 # default == kind 6
        assert np.array_equal(out[0, prec][0], out[6, prec][0])   # Newton table incl. residuals
        assert np.array_equal(out[0, prec][1], out[6, prec][1])


def test_kernel_kinds_keep_cg_counts_and_displacements(libs):
    """Kinds 3 / 6 (= the default) against the single-ring kernel in whole coupled runs: y = A x is
    bitwise the same and the dot products are kernel-independent chunked reductions, so the CG
    history must agree (iteration counts exactly, displacements to rounding)."""
    capi, solvers, mg = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                  max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    out = {}
    for kind in (5, 3, 6, 0):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_SPMV_KERNEL, kind)
        part = solvers.FakeParticipant(3, 2, p.delta_t, traction, 2)
        solid = solvers.Solid(prob, part, handle=H.fine)
        solid.run()
        out[kind] = (np.array([r for rows in solid.history for r in rows]),
                     np.array([d for (w, it, d) in part.written]))
        H.close()
    for kind in (3, 6, 0):
        assert out[kind][0].shape == out[5][0].shape                   # same Newton counts
        assert np.array_equal(out[kind][0][:, 0], out[5][0][:, 0])     # same CG iteration counts
        assert np.abs(out[kind][1] - out[5][1]).max() <= 1e-11 * np.abs(out[5][1]).max()


def test_two_ring_kernel_linear_model(libs):
    capi, solvers, mg = libs
    p = lin_params(poly_degree=1, type_lin="CG", max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[6, 20, 6])
    buf = np.tile([300.0, -100.0, 50.0], prob.n_iface_nodes)
    out = {}
    for kind in (5, 0, 3, 6):
        h = capi.Handle(prob)
        h.set_option(capi.OPT_SPMV_KERNEL, kind)
        part = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: buf)
        ed = solvers.ElastoDynamics(prob, part, handle=h)
        ed.run()
        out[kind] = (np.array(ed.history), part.written[-1][2])
        h.close()
    assert np.array_equal(out[0][0], out[6][0]) and np.array_equal(out[0][1], out[6][1])
    for kind in (3, 6):
        assert np.array_equal(out[kind][0][:, 0], out[5][0][:, 0])     # CG iteration counts
        assert np.abs(out[kind][1] - out[5][1]).max() <= 1e-11 * np.abs(out[5][1]).max()
