"""GPU tests of the matrix-free tangent operator (GF_OPT_OPERATOR = 1, csrc/matfree.cu): it must
apply the same bilinear form the reference assembles (nonlinear_elasticity.cc:1011-1023) with the
Dirichlet treatment of distribute_local_to_global (:769-773)."""
import numpy as np
import pytest

from helpers import nl_params, rel_err, smooth_field
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import build, capi, multigrid, solvers
    from oracle import oracle_py as orc
    build.build_cuda()
    capi.lib()
    return capi, solvers, multigrid, orc


@pytest.mark.parametrize("degree,reps,numbering", [
    (2, [3, 4, 2], "cellwise"),
    (2, [2, 5, 3], "component_wise"),
    (1, [4, 5, 3], "lexicographic"),
])
def test_matrix_free_operator_matches_assembled_tangent(libs, degree, reps, numbering):
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=degree, body_force=(0.0, -9.81, 2.0))
    prob = make_problem(p, 3, reps=reps, numbering=numbering)
    h = capi.Handle(prob)
    rng = np.random.RandomState(11)
    u = smooth_field(prob, 0.02, seed=2)
    a = smooth_field(prob, 30.0, seed=3)
    stress = np.zeros(prob.n_dofs)
    stress[prob.iface_dofs.reshape(-1)] = 800.0
    x = rng.uniform(-1, 1, prob.n_dofs)          # non-zero on constrained dofs on purpose
    out = {}
    for op in (0, 1):
        h.set_option(capi.OPT_OPERATOR, op)
        h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
        h.set_vector(capi.NL_VELOCITY_OLD, a * 1e-3)
        h.set_vector(capi.NL_ACCELERATION_OLD, a)
        h.set_vector(capi.NL_EXTERNAL_STRESS, stress)
        h.nl_begin_step()
        res = h.nl_newton_assemble()
        h.set_vector(capi.VEC_SCRATCH0, x)
        h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
        out[op] = (res, h.get_vector(capi.NL_SYSTEM_RHS), h.get_vector(capi.VEC_SCRATCH1))
    assert abs(out[0][0] - out[1][0]) <= 1e-12 * out[0][0]
    assert rel_err(out[1][1], out[0][1]) < 1e-12
    assert rel_err(out[1][2], out[0][2]) < 1e-12
    with pytest.raises(capi.GraftError):
        h.export_csr(capi.MAT_TANGENT)            # nothing assembled in matrix-free mode
    h.close()


def test_matrix_free_cg_newton_counts_and_displacement_match_oracle(libs):
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="Direct", delta_t=0.01)
    prob = make_problem(p, 3, reps=[3, 6, 2])
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.03), n)
    h = capi.Handle(prob)
    h.set_option(capi.OPT_OPERATOR, 1)
    part = solvers.FakeParticipant(3, 3, p.delta_t, traction)
    solid = solvers.Solid(prob, part, handle=h)
    solid.run()
    o = orc.Oracle(prob)
    counts = []
    for w in range(3):
        o.format_precice_to_deal(traction((w + 1) * p.delta_t, 0), orc.NL_EXTERNAL_STRESS)
        k, _ = o.nl_timestep()
        counts.append(k)
        ref = o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT)
        assert rel_err(part.written[w][2], ref) < 1e-8
    assert [len(r) for r in solid.history] == counts
    h.close()


def test_matrix_free_finest_level_inside_multigrid(libs):
    capi, solvers, mg, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01,
                  max_iterations_lin=1.0)
    prob = make_problem(p, 3, reps=[4, 16, 4], numbering="lexicographic")
    n = prob.n_iface_nodes
    buf = np.tile([2000.0, 0.0, 0.0], n)
    out = {}
    for op in (0, 1):
        H = mg.Hierarchy(prob)
        H.fine.set_option(capi.OPT_OPERATOR, op)
        part = solvers.FakeParticipant(3, 2, p.delta_t, lambda t, it: buf)
        solid = solvers.Solid(prob, part, handle=H.fine)
        solid.run()
        out[op] = ([[r[0] for r in rows] for rows in solid.history], part.written[-1][2])
        H.close()
    assert [len(r) for r in out[0][0]] == [len(r) for r in out[1][0]]
    assert max(max(r) for r in out[1][0]) <= max(max(r) for r in out[0][0]) + 2
    assert rel_err(out[1][1], out[0][1]) < 1e-7


def test_matrix_free_rejected_where_unsupported(libs):
    capi, solvers, mg, orc = libs
    h = capi.Handle(make_problem(nl_params(poly_degree=2), 2, reps=[2, 4]))
    with pytest.raises(capi.GraftError) as e:
        h.set_option(capi.OPT_OPERATOR, 1)
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    h.close()
