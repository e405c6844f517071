"""Polynomial degree >= 3 on the device, against the oracle. The reference's shipped parameter
files use it: parameters.prm:21 (degree 3, linear model, FSI3, -DDIM=2 default of CMakeLists.txt:15)
and source/nonlinear_elasticity/nonlinear_elasticity.prm:24 (degree 4, neo-Hookean, FSI3).
Device side: FE_Q tables on Gauss-Lobatto points and the FESystem local numbering for any degree
(csrc/fe_basis.h), the generic-degree cell / face kernels (csrc/assemble_nl_generic.cuh; their CPU
emulation is tests/test_cuda_emulation.py), the runtime-degree linear kernels, block-Jacobi CG.
The file sorts last on purpose: these kernels were written in a session without GPU access, so the
round-end run of this file is their first execution on hardware."""
import numpy as np
import pytest

from helpers import lin_params, nl_params, rel_err
from dealii_adapter_b200.problem import SolverParameters, make_problem
from test_gpu_parity import (assert_matrix_close, nl_state, run_nonlinear, run_oracle_nonlinear)

pytestmark = pytest.mark.gpu

# The driver runs `pytest -x`: one failing test would hide every later one. These tests are the
# FIRST execution of new kernels on hardware, so each of them records a failure (reported as
# xfailed, with the reason) and lets the run continue; the last test of the file then fails with
# the complete list. Nothing is hidden: a green file means every test below passed.
_FIRST_RUN_FAILURES = []


def first_run(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except (pytest.skip.Exception, KeyboardInterrupt):
            raise
        except BaseException as exc:     # assertion, GraftError, subprocess failure ...
            _FIRST_RUN_FAILURES.append("%s%s: %s" % (fn.__name__, tuple(kwargs.get(k) for k in
                                       ("dim", "degree", "case") if k in kwargs), repr(exc)[:400]))
            pytest.xfail("first hardware run failed: %s" % repr(exc)[:300])
    return wrapper


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import capi, solvers
    from oracle import oracle_py
    native_libs.build_cuda()
    capi.lib()
    return capi, solvers, oracle_py


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (2, 3, [3, 4], "cellwise"),
    (2, 4, [2, 3], "component_wise"),
    (2, 5, [2, 2], "lexicographic"),
    (3, 3, [2, 2, 1], "cellwise"),
    (3, 3, [1, 2, 2], "lexicographic"),
])
@first_run
def test_nonlinear_tangent_and_residual_match_oracle(libs, dim, degree, reps, numbering):
    capi, solvers, orc = libs
    p = nl_params(poly_degree=degree, body_force=(0.3, -9.81, 0.2 if dim == 3 else 0.0))
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    u, du, v_old, a_old, traction = nl_state(prob)
    o = orc.Oracle(prob)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    o.set(orc.NL_SOLUTION_DELTA, du)
    o.set(orc.NL_VELOCITY_OLD, v_old)
    o.set(orc.NL_ACCELERATION_OLD, a_old)
    o.format_precice_to_deal(traction, orc.NL_EXTERNAL_STRESS)
    o.nl_update_acceleration()
    o.nl_assemble_system()
    h = capi.Handle(prob)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.set_vector(capi.NL_SOLUTION_DELTA, du)
    h.set_vector(capi.NL_VELOCITY_OLD, v_old)
    h.set_vector(capi.NL_ACCELERATION_OLD, a_old)
    h.set_traction(traction)
    res = h.nl_newton_assemble()
    rowptr_o, col_o = o.pattern()
    assert_matrix_close(*h.export_csr(capi.MAT_TANGENT), rowptr_o, col_o, o.values(orc.MAT_TANGENT))
    rhs_o = o.get(orc.NL_SYSTEM_RHS)
    assert rel_err(h.get_vector(capi.NL_SYSTEM_RHS), rhs_o) < 1e-12
    assert abs(res - o.nl_error_residual()) <= 1e-12 * o.nl_error_residual()
    v1 = h.export_csr(capi.MAT_TANGENT)[2]
    r1 = h.get_vector(capi.NL_SYSTEM_RHS)
    h.nl_newton_assemble()                       # bitwise reproducible
    assert np.array_equal(v1, h.export_csr(capi.MAT_TANGENT)[2])
    assert np.array_equal(r1, h.get_vector(capi.NL_SYSTEM_RHS))
    # operator application of the assembled tangent (rows of up to (2p+1)^dim blocks)
    x = np.random.RandomState(3).uniform(-1, 1, prob.n_dofs)
    h.set_vector(capi.VEC_SCRATCH0, x)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y_o = o.vmult(orc.MAT_TANGENT, x)
    assert rel_err(h.get_vector(capi.VEC_SCRATCH1), y_o) < 1e-13
    h.close()


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (2, 3, [3, 6], "cellwise"),
    (2, 4, [2, 4], "lexicographic"),
    (3, 3, [2, 2, 1], "component_wise"),
])
@first_run
def test_linear_matrices_and_steps_match_oracle(libs, dim, degree, reps, numbering):
    capi, solvers, orc = libs
    p = lin_params(poly_degree=degree, body_force=(0.0, -9.81, 0.0), type_lin="CG")
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    h = capi.Handle(prob)
    h.lin_assemble_once()
    rowptr_o, col_o = o.pattern()
    assert_matrix_close(*h.export_csr(capi.MAT_STIFFNESS), rowptr_o, col_o, o.values(orc.MAT_STIFFNESS))
    assert_matrix_close(*h.export_csr(capi.MAT_MASS), rowptr_o, col_o, o.values(orc.MAT_MASS))
    assert rel_err(h.get_vector(capi.LIN_BODY_FORCE), o.get(orc.LIN_BODY_FORCE)) < 1e-12
    n = prob.n_iface_nodes
    load = np.array([300.0, -100.0, 50.0][:dim])
    for step in range(2):
        buf = np.tile(load * (step + 1), n)
        o.format_precice_to_deal(buf, orc.LIN_STRESS)
        o.lin_step()
        h.set_traction(buf)
        it_g, res_g = h.lin_step(0, 8.0)
        if step == 0:
            assert_matrix_close(*h.export_csr(capi.MAT_SYSTEM), rowptr_o, col_o,
                                o.values(orc.MAT_SYSTEM))
        assert res_g <= 1e-10 and it_g > 0
        assert rel_err(h.get_vector(capi.LIN_OLD_STRESS), o.get(orc.LIN_OLD_STRESS)) < 1e-12
        assert np.abs(h.get_vector(capi.LIN_VELOCITY) - o.get(orc.LIN_VELOCITY)).max() < 1e-7
        assert np.abs(h.get_vector(capi.LIN_DISPLACEMENT) - o.get(orc.LIN_DISPLACEMENT)).max() < 1e-9
    h.close()


@first_run
def test_shipped_default_linear_degree3_fsi3_direct(libs):
    """parameters.prm as shipped: linear model, FSI3, degree 3, Solver type = Direct, dt 0.005,
    nu 0.4, mu 0.5e6, rho 1000 (lines 9, 21, 25-31, 40-43, 63); 2D build."""
    capi, solvers, orc = libs
    p = SolverParameters(model="linear", type_lin="Direct", poly_degree=3, scenario="FSI3",
                         delta_t=0.005, mu=0.5e6, nu=0.4, rho=1000.0, max_iterations_lin=1.0)
    prob = make_problem(p, 2)                    # 18 x 3 cells, 55 x 10 nodes
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([0.0, -20.0]), n)
    part = solvers.FakeParticipant(2, 4, p.delta_t, traction)
    ed = solvers.ElastoDynamics(prob, part)
    ed.run()
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    for step in range(4):
        o.format_precice_to_deal(traction(0, 0), orc.LIN_STRESS)
        o.lin_step()
        assert rel_err(part.written[step][2], o.format_deal_to_precice(orc.LIN_DISPLACEMENT)) < 1e-8
    ed.handle.close()


@first_run
def test_shipped_default_nonlinear_degree4_fsi3_direct(libs):
    """source/nonlinear_elasticity/nonlinear_elasticity.prm as shipped: neo-Hookean, FSI3, degree 4,
    Solver type = Direct (lines 24, 46-49, 65): identical Newton counts, interface displacement
    1e-8, implicit coupling with a checkpoint restore in between."""
    capi, solvers, orc = libs
    p = nl_params(poly_degree=4, scenario="FSI3", type_lin="Direct", delta_t=0.01)
    prob = make_problem(p, 2)                    # 18 x 3 cells
    n = prob.n_iface_nodes
    relax = [0.7, 1.0]

    def traction(t, it):
        return np.tile(np.array([0.0, -1500.0]) * min(1.0, t / 0.02) * relax[it], n)

    solid, part = run_nonlinear(libs, prob, 3, traction, n_sub=2)
    o, counts, written = run_oracle_nonlinear(orc, prob, 3, traction, n_sub=2)
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-8
    solid.handle.close()


@first_run
def test_cg_path_degree3_3d(libs):
    """'Solver type = CG' on 3D Q3 hexahedra (rows of up to 343 node blocks): device block-Jacobi
    CG vs the oracle's SSOR-CG, same Newton counts, displacements to the inexact-Newton level."""
    capi, solvers, orc = libs
    p = nl_params(poly_degree=3, scenario="PF", type_lin="CG", delta_t=0.01)
    prob = make_problem(p, 3, reps=[1, 4, 1])
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1200.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    solid, part = run_nonlinear(libs, prob, 2, traction)
    o, counts, written = run_oracle_nonlinear(orc, prob, 2, traction)
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-6
    solid.handle.close()


@pytest.mark.parametrize("dim,degree,reps", [(2, 3, [3, 2]), (3, 3, [1, 2, 1])])
@first_run
def test_output_fields_match_oracle(libs, dim, degree, reps):
    """gf_postprocess (DataOut patches through MappingQEulerian + Postprocessor) at degree 3."""
    capi, solvers, orc = libs
    prob = make_problem(nl_params(poly_degree=degree), dim, reps=reps)
    u, _, _, _, _ = nl_state(prob)
    o = orc.Oracle(prob)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    pts_o, fld_o = o.postprocess(orc.NL_TOTAL_DISPLACEMENT)
    h = capi.Handle(prob)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    fld = h.postprocess(capi.NL_TOTAL_DISPLACEMENT)
    assert np.abs(fld - fld_o).max() <= 1e-12 * max(1.0, np.abs(fld_o).max())
    h.close()


@first_run
def test_multigrid_and_matrix_free_are_refused_above_degree_2(libs):
    capi, solvers, orc = libs
    from dealii_adapter_b200 import multigrid
    prob = make_problem(nl_params(poly_degree=3), 3, reps=[2, 2, 2])
    with pytest.raises(capi.GraftError) as e:
        multigrid.Hierarchy(prob)
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    h = capi.Handle(prob)
    with pytest.raises(capi.GraftError) as e:
        h.set_option(capi.OPT_OPERATOR, 1)
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    h.close()
    with pytest.raises(capi.GraftError) as e:
        capi.Handle(make_problem(nl_params(poly_degree=4), 3, reps=[1, 1, 1]))
    assert e.value.code == capi.GF_ERR_UNSUPPORTED


# The key/value content of the reference's shipped parameters.prm (comments dropped; a CPU test
# below compares it with the real file where /root/reference exists). Linear model, FSI3,
# degree 3, Direct; the driver is built with -DDIM=2 like the reference's default.
SHIPPED_PARAMETERS_PRM = """subsection Time
  set End time              = 10
  set Time step size        = 0.005
  set Output interval       = 10
  set Output folder   = dealii-output
end
subsection Discretization
  set Polynomial degree   = 3
end
subsection System properties
  set Poisson's ratio = 0.4
  set Shear modulus   = 0.5e6
  set rho\t      = 1000
  set body forces     = 0.0,0.0,0.0
end
subsection Solver
  set Model                     = linear
  set Solver type               = Direct
  set Max iteration multiplier  = 1
  set Residual                  = 1e-6
  set Max iterations Newton-Raphson = 10
  set Tolerance displacement        = 1.0e-6
  set Tolerance force               = 1.0e-9
end
subsection precice configuration
  set Scenario            = FSI3
  set precice config-file = precice-config.xml
  set Participant name    = Solid
  set Mesh name           = Solid-Mesh
  set Read data name      = Stress
  set Write data name     = Displacement
end
"""


@first_run
def test_shipped_parameter_file_runs_unchanged_through_the_cpp_driver(tmp_path, native_libs):
    """elasticity_2d with the reference's own parameters.prm (degree 3): the scripted participant
    is configured by a file of the name the prm asks for; 5 time windows of 0.005; watch point vs
    the oracle."""
    import subprocess
    from dealii_adapter_b200 import build
    from oracle import oracle_py as orc
    exe = build.build_elasticity()[0]
    (tmp_path / "parameters.prm").write_text(SHIPPED_PARAMETERS_PRM)
    (tmp_path / "precice-config.xml").write_text(
        "dimensions = 2\ntime-window-size = 0.005\nmax-time-windows = 5\nsub-iterations = 1\n"
        "traction = 0.0,-20.0\nramp-time = 0.0\nwatch-point = 0.6,0.2\n"
        "watch-point-file = watchpoint.log\n")
    import os
    r = subprocess.run([exe, "parameters.prm"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, GF_DIRECT_SOLVER="auto"))
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    assert "Polynomial degree: 3" in r.stdout
    assert "Direct solver: band Cholesky on the device" in r.stdout
    vtk = (tmp_path / "dealii-output" / "solution-000.vtk").read_text()
    assert "POINTS %d double" % (54 * 16) in vtk          # one order-3 Lagrange quad per cell
    log = np.loadtxt(tmp_path / "watchpoint.log")
    assert log.shape == (5, 6)
    p = SolverParameters(model="linear", type_lin="Direct", poly_degree=3, scenario="FSI3",
                         delta_t=0.005, mu=0.5e6, nu=0.4, rho=1000.0, max_iterations_lin=1.0)
    prob = make_problem(p, 2, numbering="component_wise")
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    pos = prob.interface_positions().reshape(-1, 2)
    k = np.argmin(((pos - np.array([0.6, 0.2])) ** 2).sum(axis=1))
    assert np.allclose(log[0, 2:4], pos[k])
    for step in range(5):
        o.format_precice_to_deal(np.tile(np.array([0.0, -20.0]), prob.n_iface_nodes), orc.LIN_STRESS)
        o.lin_step()
        d = o.format_deal_to_precice(orc.LIN_DISPLACEMENT).reshape(-1, 2)[k]
        assert rel_err(log[step, 4:6], d) < 1e-8


# ------------------------------------------------------------------------------------------------
# 'Solver type = Direct' (the shipped default, parameters.prm:43): band Cholesky on the device
# (csrc/direct_band.cuh; CPU emulation: tests/test_cuda_emulation.py). Also first run on hardware.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,degree,scenario,reps,load", [
    (2, 2, "FSI3", [18, 3], (0.0, -1500.0)),
    (3, 2, "PF", [3, 6, 2], (1500.0, 0.0, 0.0)),
    (2, 3, "FSI3", [18, 3], (0.0, -1500.0)),
])
@first_run
def test_direct_solver_band_cholesky_nonlinear(libs, dim, degree, scenario, reps, load):
    capi, solvers, orc = libs
    p = nl_params(poly_degree=degree, scenario=scenario, type_lin="Direct", delta_t=0.01)
    prob = make_problem(p, dim, reps=reps)
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array(load) * min(1.0, t / 0.03), n)
    part = solvers.FakeParticipant(dim, 3, p.delta_t, traction)
    solid = solvers.Solid(prob, part)
    solid.handle.set_option(capi.OPT_DIRECT_SOLVER, 0)         # auto: band Cholesky where it fits
    solid.run()
    n_solves, w, res = solid.handle.direct_info()
    assert n_solves == sum(len(r) for r in solid.history)      # every solve by the factorisation
    assert res <= 1e-9 and 0 < w < prob.n_dofs
    o, counts, written = run_oracle_nonlinear(orc, prob, 3, traction)
    assert [len(r) for r in solid.history] == counts
    for (w_, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-8
    # the CG stand-in (GF_OPT_DIRECT_SOLVER = 2) gives the same displacements
    part2 = solvers.FakeParticipant(dim, 3, p.delta_t, traction)
    solid2 = solvers.Solid(prob, part2)
    solid2.handle.set_option(capi.OPT_DIRECT_SOLVER, 2)
    solid2.run()
    with pytest.raises(capi.GraftError) as e:
        solid2.handle.direct_info()
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    for a, b in zip(part.written, part2.written):
        assert rel_err(a[2], b[2]) < 1e-9
    solid.handle.close()
    solid2.handle.close()


@first_run
def test_direct_solver_linear_factorises_once_and_falls_back_beyond_the_budget(libs, monkeypatch):
    capi, solvers, orc = libs
    p = lin_params(poly_degree=2, type_lin="Direct")
    prob = make_problem(p, 2, reps=[6, 36])
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([50.0, 0.0]), n)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    refs = []
    for step in range(4):
        o.format_precice_to_deal(traction(0, 0), orc.LIN_STRESS)
        o.lin_step()
        refs.append(o.format_deal_to_precice(orc.LIN_DISPLACEMENT))
    part = solvers.FakeParticipant(2, 4, p.delta_t, traction)
    ed = solvers.ElastoDynamics(prob, part)
    ed.handle.set_option(capi.OPT_DIRECT_SOLVER, 0)
    ed.run()
    n_solves, w, res = ed.handle.direct_info()
    assert n_solves == 4 and res <= 1e-9
    for step in range(4):
        assert rel_err(part.written[step][2], refs[step]) < 1e-8
    ed.handle.close()
    # band beyond the memory budget: the handle answers with the CG stand-in, same results
    monkeypatch.setenv("GF_DIRECT_BUDGET_MB", "0.01")
    part = solvers.FakeParticipant(2, 4, p.delta_t, traction)
    ed = solvers.ElastoDynamics(prob, part)
    ed.handle.set_option(capi.OPT_DIRECT_SOLVER, 0)
    ed.run()
    with pytest.raises(capi.GraftError) as e:
        ed.handle.direct_info()
    assert e.value.code == capi.GF_ERR_UNSUPPORTED and "budget" in str(e.value)
    for step in range(4):
        assert rel_err(part.written[step][2], refs[step]) < 1e-8
    # ... and GF_OPT_DIRECT_SOLVER = 1 insists
    ed.handle.set_option(capi.OPT_DIRECT_SOLVER, 1)
    with pytest.raises(capi.GraftError) as e:
        ed.handle.lin_step(1, 1.0)
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    ed.handle.close()


# ------------------------------------------------------------------------------------------------
# the reference's OWN assembly blocks at degree 3 / 4 (tests/golden/reference_vectors.npz, cases
# added with the high-degree support; CPU: tests/test_reference_pins.py pins the oracle with them)
# ------------------------------------------------------------------------------------------------
def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                "reference_vectors.npz"))


def _to_global(prob, v_local):
    out = np.zeros(prob.n_dofs)
    out[prob.mesh.cell_dofs.reshape(-1)] = v_local
    return out


@pytest.mark.parametrize("case", [5, 6, 7, 8, 9, 10])
@first_run
def test_device_cell_assembly_equals_the_reference_assembly_block(libs, case):
    """cases 5-7: degree 3 / 4 on Cartesian cells; 8-10: GENERAL (non-affine) cells of degree 2, 1, 3"""
    import scipy.sparse as sp
    capi, solvers, orc = libs
    ref = _golden()
    meta = ref["asm%d_meta" % case]
    dim, degree = int(meta[0]), int(meta[1])
    h, body_force = meta[2:5], meta[5:8]
    mu, nu, rho, beta, dt = meta[8:13]
    verts = ref["asm%d_verts" % case] if "asm%d_verts" % case in ref.files else None
    assert (degree >= 3 or verts is not None) and sorted(ref["asm%d_faces" % case].tolist()) == [0, 1, 3]
    p = nl_params(poly_degree=degree, scenario="PF", mu=mu, nu=nu, rho=rho, beta=beta, delta_t=dt,
                  body_force=tuple(body_force))
    prob = make_problem(p, dim, reps=[1] * dim, numbering="cellwise",
                        box=([0.0] * dim, list(h[:dim]) if verts is None else [1.0] * dim))
    if verts is not None:
        prob.mesh.cell_vertices = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1)
    prob.constrained = np.zeros_like(prob.constrained)
    cd = prob.mesh.cell_dofs.reshape(-1)       # local dof i of the reference block = global cd[i]
    hd = capi.Handle(prob)
    hd.set_vector(capi.NL_TOTAL_DISPLACEMENT, _to_global(prob, ref["asm%d_u" % case]))
    hd.set_vector(capi.NL_EXTERNAL_STRESS, _to_global(prob, ref["asm%d_stress" % case]))
    alpha_3 = (1 - 2 * beta) / (2 * beta)      # update_acceleration (:592-599) with v_old = 0
    hd.set_vector(capi.NL_ACCELERATION_OLD, _to_global(prob, -ref["asm%d_acc" % case] / alpha_3))
    hd.nl_begin_step()
    hd.nl_newton_assemble()
    rowptr, col, val = hd.export_csr(capi.MAT_TANGENT)
    K = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs)).toarray()[np.ix_(cd, cd)]
    r = hd.get_vector(capi.NL_SYSTEM_RHS)[cd]
    K_ref, r_ref = ref["asm%d_K" % case], ref["asm%d_r" % case]
    assert np.abs(K - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    assert np.abs(r - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
    hd.close()


@pytest.mark.parametrize("case", [4, 5, 6])
@first_run
def test_device_linear_stiffness_and_loading_equal_the_reference_loops(libs, case):
    import scipy.sparse as sp
    capi, solvers, orc = libs
    ref = _golden()
    meta = ref["lin%d_meta" % case]
    dim, degree, h, mu, nu = int(meta[0]), int(meta[1]), meta[2:5], meta[5], meta[6]
    assert degree >= 3
    p = lin_params(poly_degree=degree, scenario="PF", mu=mu, nu=nu, type_lin="CG")
    prob = make_problem(p, dim, reps=[1] * dim, numbering="cellwise",
                        box=([0.0] * dim, list(h[:dim])))
    prob.constrained = np.zeros_like(prob.constrained)
    cd = prob.mesh.cell_dofs.reshape(-1)
    hd = capi.Handle(prob)
    hd.lin_assemble_once()
    rowptr, col, val = hd.export_csr(capi.MAT_STIFFNESS)
    K = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs)).toarray()[np.ix_(cd, cd)]
    K_ref, F_ref = ref["lin%d_K" % case], ref["lin%d_F" % case]
    assert np.abs(K - K_ref).max() <= 1e-12 * np.abs(K_ref).max()
    hd.set_vector(capi.LIN_STRESS, _to_global(prob, ref["lin%d_stress" % case]))
    hd.lin_step(0, 20.0)                      # old_stress <- the consistent loading (:405-409)
    F = hd.get_vector(capi.LIN_OLD_STRESS)[cd]
    assert np.abs(F - F_ref).max() <= 1e-12 * np.abs(F_ref).max()
    hd.close()


# ------------------------------------------------------------------------------------------------
# general (non-affine) cells: what a real deal.II host hands over once the mesh is not a box.
# Device: Jacobians per quadrature point (assemble_nl_generic.cuh with AFFINE = false,
# assemble_general.cuh); CPU emulation of the same kernels: tests/test_cuda_emulation.py.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (2, 1, [4, 5], "cellwise"),
    (2, 2, [3, 4], "component_wise"),
    (2, 3, [3, 3], "cellwise"),
    (3, 1, [3, 3, 2], "lexicographic"),
    (3, 2, [2, 3, 2], "cellwise"),
])
@first_run
def test_distorted_mesh_nonlinear_tangent_residual_and_output(libs, dim, degree, reps, numbering):
    from helpers import distort_mesh
    capi, solvers, orc = libs
    p = nl_params(poly_degree=degree, body_force=(0.3, -9.81, 0.2 if dim == 3 else 0.0))
    prob = distort_mesh(make_problem(p, dim, reps=reps, numbering=numbering), 0.12, seed=degree)
    u, du, v_old, a_old, traction = nl_state(prob)
    o = orc.Oracle(prob)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    o.set(orc.NL_SOLUTION_DELTA, du)
    o.set(orc.NL_VELOCITY_OLD, v_old)
    o.set(orc.NL_ACCELERATION_OLD, a_old)
    o.format_precice_to_deal(traction, orc.NL_EXTERNAL_STRESS)
    o.nl_update_acceleration()
    o.nl_assemble_system()
    h = capi.Handle(prob)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.set_vector(capi.NL_SOLUTION_DELTA, du)
    h.set_vector(capi.NL_VELOCITY_OLD, v_old)
    h.set_vector(capi.NL_ACCELERATION_OLD, a_old)
    h.set_traction(traction)
    res = h.nl_newton_assemble()
    rowptr_o, col_o = o.pattern()
    assert_matrix_close(*h.export_csr(capi.MAT_TANGENT), rowptr_o, col_o, o.values(orc.MAT_TANGENT))
    assert rel_err(h.get_vector(capi.NL_SYSTEM_RHS), o.get(orc.NL_SYSTEM_RHS)) < 1e-12
    assert abs(res - o.nl_error_residual()) <= 1e-12 * o.nl_error_residual()
    _, fld_o = o.postprocess(orc.NL_TOTAL_DISPLACEMENT)
    fld = h.postprocess(capi.NL_TOTAL_DISPLACEMENT)
    assert np.abs(fld - fld_o).max() <= 1e-12 * max(1.0, np.abs(fld_o).max())
    h.close()


@pytest.mark.parametrize("dim,degree,reps", [(2, 2, [3, 8]), (3, 1, [3, 4, 2]), (2, 3, [2, 4])])
@first_run
def test_distorted_mesh_linear_matrices_and_steps(libs, dim, degree, reps):
    from helpers import distort_mesh
    capi, solvers, orc = libs
    p = lin_params(poly_degree=degree, body_force=(0.0, -9.81, 0.0), type_lin="CG")
    prob = distort_mesh(make_problem(p, dim, reps=reps), 0.12, seed=7)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    h = capi.Handle(prob)
    h.lin_assemble_once()
    rowptr_o, col_o = o.pattern()
    assert_matrix_close(*h.export_csr(capi.MAT_STIFFNESS), rowptr_o, col_o, o.values(orc.MAT_STIFFNESS))
    assert_matrix_close(*h.export_csr(capi.MAT_MASS), rowptr_o, col_o, o.values(orc.MAT_MASS))
    assert rel_err(h.get_vector(capi.LIN_BODY_FORCE), o.get(orc.LIN_BODY_FORCE)) < 1e-12
    n = prob.n_iface_nodes
    load = np.array([300.0, -100.0, 50.0][:dim])
    for step in range(2):
        buf = np.tile(load * (step + 1), n)
        o.format_precice_to_deal(buf, orc.LIN_STRESS)
        o.lin_step()
        h.set_traction(buf)
        it_g, res_g = h.lin_step(0, 8.0)
        assert res_g <= 1e-10 and it_g > 0
        assert rel_err(h.get_vector(capi.LIN_OLD_STRESS), o.get(orc.LIN_OLD_STRESS)) < 1e-12
        assert np.abs(h.get_vector(capi.LIN_DISPLACEMENT) - o.get(orc.LIN_DISPLACEMENT)).max() < 1e-9
    h.close()


@first_run
def test_distorted_mesh_coupled_run_and_multigrid_refusal(libs):
    from helpers import distort_mesh
    from dealii_adapter_b200 import multigrid
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="Direct", delta_t=0.01)
    prob = distort_mesh(make_problem(p, 3, reps=[2, 6, 2]), 0.1, seed=4)
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1500.0, 0.0, 0.0]) * min(1.0, t / 0.03), n)
    solid, part = run_nonlinear(libs, prob, 3, traction)
    o, counts, written = run_oracle_nonlinear(orc, prob, 3, traction)
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-8
    solid.handle.close()
    with pytest.raises(capi.GraftError) as e:
        multigrid.Hierarchy(prob)
    assert e.value.code == capi.GF_ERR_UNSUPPORTED



@first_run
def test_partitioned_run_at_degree_3_reproduces_the_single_rank_run(native_libs, tmp_path):
    """Slab partition at degree 3 (node planes on Gauss-Lobatto coordinates, halo of p = 3 planes):
    2 ranks on one device through the NCCL-free bootstrap, block-Jacobi CG, against one rank -
    identical Newton and CG counts, bitwise displacements (partition-independent reductions)."""
    native_libs.build_cuda()
    import mgpu_worker as w
    from test_gpu_multirank import _compare, _spawn
    ref = {}
    for name in w.EXTRA_CASES:
        hist, written, levels = w.run_case(name, 1, 0, 0, None)
        ref[name] = {"written": written, "history": hist, "levels": levels}
    got = _spawn(2, "ipc", str(tmp_path / "q3.pkl"), cases=",".join(w.EXTRA_CASES))
    for name in w.EXTRA_CASES:
        _compare(ref, got, name, True)


# ------------------------------------------------------------------------------------------------
# hanging-node constraints (gf_desc.line_*): make_hanging_node_constraints / condense / distribute
# of linear_elasticity.cc:196-207,355,422,571 and the AffineConstraints object of the nonlinear
# solver. Mesh: tests/helpers.py::hanging_node_problem; reference: condensation by definition
# (helpers.reference_*; tests/test_hanging_node_reference.py checks it on the CPU).
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
@first_run
def test_hanging_nodes_linear_steps(libs, dim, degree):
    from helpers import constraint_matrix, hanging_node_problem, reference_linear_steps
    capi, solvers, orc = libs
    p = lin_params(poly_degree=degree, type_lin="CG", body_force=(0.0, -9.81, 0.0),
                   max_iterations_lin=20.0)
    prob = hanging_node_problem(p, degree, dim)
    n = prob.n_iface_nodes
    bufs = [np.tile([40.0 * (k + 1), -200.0, 30.0][:dim], n) for k in range(3)]
    ref = reference_linear_steps(orc, prob, bufs)
    part = solvers.FakeParticipant(dim, 3, p.delta_t, lambda t, it: bufs[min(2, int(round(t / p.delta_t)) - 1)])
    ed = solvers.ElastoDynamics(prob, part)
    ed.run()
    d = ed.handle.get_vector(capi.LIN_DISPLACEMENT)
    assert rel_err(d, ref[-1]) < 1e-6
    dof = prob.extra["constraint_lines"][0]
    m = d.copy()
    m[dof] = 0.0
    assert np.abs(constraint_matrix(prob) @ m - d).max() <= 1e-14 * np.abs(d).max()   # conforming
    for k in range(3):
        assert rel_err(part.written[k][2], ref[k][prob.iface_dofs.T].reshape(-1)) < 1e-6
    ed.handle.close()
    # 'Solver type = Direct' (linear_elasticity.cc:556-563): the factor of the constant system
    # matrix, computed once, preconditions the CG on the condensed operator
    pd = lin_params(poly_degree=degree, type_lin="Direct", body_force=(0.0, -9.81, 0.0))
    probd = hanging_node_problem(pd, degree, dim)
    partd = solvers.FakeParticipant(dim, 3, pd.delta_t, lambda t, it: bufs[min(2, int(round(t / pd.delta_t)) - 1)])
    edd = solvers.ElastoDynamics(probd, partd)
    edd.handle.set_option(capi.OPT_DIRECT_SOLVER, 0)
    edd.run()
    assert rel_err(edd.handle.get_vector(capi.LIN_DISPLACEMENT), ref[-1]) < 1e-9
    n_solves, half_bw, res = edd.handle.direct_info()
    assert n_solves == 3 and half_bw > 0 and res <= 1e-10
    edd.handle.close()


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
@first_run
def test_hanging_nodes_nonlinear_step_direct_and_cg(libs, dim, degree):
    from helpers import constraint_matrix, hanging_node_problem, reference_nonlinear_step
    capi, solvers, orc = libs
    p = nl_params(poly_degree=degree, type_lin="Direct", scenario="PF", delta_t=0.01)
    prob = hanging_node_problem(p, degree, dim)
    buf = np.tile([0.0, -1500.0, 300.0][:dim], prob.n_iface_nodes)
    ref = reference_nonlinear_step(orc, prob, buf)
    part = solvers.FakeParticipant(dim, 1, p.delta_t, lambda t, it: buf)
    solid = solvers.Solid(prob, part)
    # auto: the band Cholesky factor of the uncondensed tangent preconditions a CG on the
    # condensed operator (csrc/cg.cu: cg_solve_direct_precond)
    solid.handle.set_option(capi.OPT_DIRECT_SOLVER, 0)
    solid.run()
    u = solid.handle.get_vector(capi.NL_TOTAL_DISPLACEMENT)
    assert 3 <= len(solid.history[0]) <= 7
    assert rel_err(u, ref) < 1e-8
    dof = prob.extra["constraint_lines"][0]
    m = u.copy()
    m[dof] = 0.0
    assert np.abs(constraint_matrix(prob) @ m - u).max() <= 1e-14 * np.abs(u).max()
    n_solves, half_bw, res = solid.handle.direct_info()
    assert n_solves == len(solid.history[0]) and half_bw > 0 and res <= 1e-10
    solid.handle.close()
    # GF_OPT_DIRECT_SOLVER = 2: the tight block-Jacobi CG on the condensed operator, same answer
    part2 = solvers.FakeParticipant(dim, 1, p.delta_t, lambda t, it: buf)
    solid2 = solvers.Solid(prob, part2)
    solid2.handle.set_option(capi.OPT_DIRECT_SOLVER, 2)
    solid2.run()
    assert len(solid2.history[0]) == len(solid.history[0])
    assert rel_err(solid2.handle.get_vector(capi.NL_TOTAL_DISPLACEMENT), u) < 1e-9
    with pytest.raises(capi.GraftError) as e:
        solid2.handle.direct_info()
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    solid2.handle.close()
    if dim == 3:
        # the matrix-free tangent takes the same condensation (operator.cu wraps either operator)
        part3 = solvers.FakeParticipant(dim, 1, p.delta_t, lambda t, it: buf)
        solid3 = solvers.Solid(prob, part3)
        solid3.handle.set_option(capi.OPT_OPERATOR, 1)
        solid3.run()
        assert len(solid3.history[0]) == len(solid.history[0])
        assert rel_err(solid3.handle.get_vector(capi.NL_TOTAL_DISPLACEMENT), u) < 1e-9
        solid3.handle.close()


def test_zzz_every_first_run_test_above_passed():
    assert not _FIRST_RUN_FAILURES, "\n".join(_FIRST_RUN_FAILURES)
