"""GPU parity at BASELINE.json configs[3]: linear_elasticity 3D cantilever, Q1 hexahedra,
128x1024x128 cells = 51,171,075 DoFs on ONE B200 (K, A = M + theta^2 dt^2 K and M resident: 73 GB).

  * ~1000 sampled rows of K, M and the constrained system matrix against the ORACLE
    (linear_elasticity.cc:248-374, 426-451; sub-mesh device of tests/sampled_rows.py run live and
    compared with the committed fixture), 1e-12 row-relative;
  * the theta-scheme right-hand side of a step with a non-trivial state at the same rows
    (assemble_rhs, :378-454) against the oracle's, 1e-12;
  * size-independent properties: TMA kernel == LDG kernel bitwise, x.Ay == y.Ax, the CG of a step
    reaches the reference's absolute tolerance 1e-10 (:542).
Separate module: the cfg3 fixtures must be released before 130 GB of HBM are taken."""
import os
import sys

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import capi, solvers
    from oracle import oracle_py
    native_libs.build_cuda()
    capi.lib()
    return capi, solvers, oracle_py


@pytest.fixture(scope="module")
def cfg4(libs):
    capi, solvers, orc = libs
    import make_sampled_rows as g
    prob = g.cfg4_problem()
    assert prob.n_dofs == 51171075 and prob.mesh.n_cells == 16777216
    h = capi.Handle(prob)
    h.lin_assemble_once()
    yield prob, h
    h.close()


def test_cfg4_sampled_rows_match_oracle(libs, cfg4):
    capi, solvers, orc = libs
    prob, h = cfg4
    import make_sampled_rows as g
    import sampled_rows as sr
    assert h.nnz() == 4099458825
    ref = g.cfg4_oracle_rows(prob)
    gold = np.load(os.path.join(HERE, "golden", "sampled_rows_cfg4.npz"))
    assert np.array_equal(ref["rows"], gold["rows"]) and np.array_equal(ref["col"], gold["col"])
    for name in ("K", "M", "A"):
        assert rel_err(ref[name], gold[name]) < 1e-13
    assert len(ref["rows"]) >= 1000
    for name, which in (("K", capi.MAT_STIFFNESS), ("M", capi.MAT_MASS), ("A", capi.MAT_SYSTEM)):
        err = sr.assert_rows_close(h.export_rows(which, ref["rows"]), ref["rowptr"], ref["col"],
                                   ref[name], 1e-12)
        print("cfg4 %s: %d rows, max row-relative error %.2e" % (name, len(ref["rows"]), err))


def test_cfg4_rhs_of_a_step_matches_oracle_at_sampled_rows(libs, cfg4):
    capi, solvers, orc = libs
    prob, h = cfg4
    import make_sampled_rows as g
    import sampled_rows as sr
    sample = sr.sample_nodes(prob, n_per_class=8, axis=1)
    rows = sample.reshape(-1)
    # cheap smooth fields (helpers.smooth_field costs ~15 s per vector on 51 M DoFs)
    x = prob.mesh.support_points
    free = prob.constrained == 0
    vel = 0.05 * np.sin(20.0 * x[:, 0] + 3.0 * x[:, 1] + 5.0 * x[:, 2]) * x[:, 1] * free
    dis = 0.004 * np.cos(11.0 * x[:, 0] - 2.0 * x[:, 1] + 7.0 * x[:, 2]) * x[:, 1] * free
    old_stress = 30.0 * np.sin(5.0 * x[:, 0] + 1.0 * x[:, 1] - 9.0 * x[:, 2]) * free
    buf = np.tile(g.CFG4_TRACTION, prob.n_iface_nodes)
    sub, dof_map, cells = sr.sub_problem(prob, sample)
    o = orc.Oracle(sub)
    o.lin_assemble_system()
    o.set(orc.LIN_VELOCITY, vel[dof_map])
    o.set(orc.LIN_DISPLACEMENT, dis[dof_map])
    o.set(orc.LIN_OLD_STRESS, old_stress[dof_map])
    o.set(orc.LIN_STRESS, sr.traction_vector(prob, buf)[dof_map])
    o.lin_assemble_rhs()
    rhs_o = o.get(orc.LIN_SYSTEM_RHS)[np.searchsorted(dof_map, rows)]
    h.set_vector(capi.LIN_VELOCITY, vel)
    h.set_vector(capi.LIN_DISPLACEMENT, dis)
    h.set_vector(capi.LIN_OLD_STRESS, old_stress)
    h.set_traction(buf)
    it, res = h.lin_step(0, prob.params.max_iterations_lin)
    assert res <= 1e-10 and it > 0
    rhs_g = h.get_vector(capi.LIN_SYSTEM_RHS)[rows]
    assert np.abs(rhs_g - rhs_o).max() <= 1e-12 * np.abs(rhs_o).max()
    assert np.array_equal(h.get_vector(capi.LIN_OLD_VELOCITY), vel)
    # update_displacement (:579-586) with the solved velocity
    dt, th = prob.params.delta_t, prob.params.theta
    v_new = h.get_vector(capi.LIN_VELOCITY)
    assert rel_err(h.get_vector(capi.LIN_DISPLACEMENT), dis + dt * th * v_new + dt * (1 - th) * vel) < 1e-14


def test_cfg4_operator_properties(libs, cfg4):
    capi, solvers, orc = libs
    prob, h = cfg4
    rng = np.random.RandomState(3)
    x, y = rng.uniform(-1, 1, prob.n_dofs), rng.uniform(-1, 1, prob.n_dofs)

    def apply(v):
        h.set_vector(capi.VEC_SCRATCH0, v)
        h.spmv(capi.MAT_SYSTEM, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
        return h.get_vector(capi.VEC_SCRATCH1)

    Ax, Ay = apply(x), apply(y)
    a, b = float(x @ Ay), float(y @ Ax)
    assert abs(a - b) <= 1e-12 * np.sqrt(float(Ax @ Ax) * float(y @ y))
    h.set_option(capi.OPT_SPMV_KERNEL, 1)
    Ax_ldg = apply(x)
    h.set_option(capi.OPT_SPMV_KERNEL, 0)
    assert np.array_equal(Ax_ldg, Ax)
