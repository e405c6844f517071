"""Parity tests proper: the CUDA path (through the C-ABI, libgraftfem.so) against the CPU oracle on
the same inputs. Tolerances are BASELINE.json's: assembled matrix entries 1e-12 relative (to the
row's largest magnitude, SURVEY 7: exactly cancelling off-diagonals make per-entry relative error
meaningless), identical Newton iteration counts, watch-point displacements 1e-8 relative."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import lin_params, nl_params, rel_err, smooth_field
from dealii_adapter_b200.problem import make_problem

pytestmark = pytest.mark.gpu

MATRIX_TOL = 1e-12


@pytest.fixture(scope="module")
def libs(native_libs):
    from dealii_adapter_b200 import capi, solvers
    from oracle import oracle_py
    native_libs.build_cuda()
    capi.lib()
    return capi, solvers, oracle_py


def assert_matrix_close(rowptr_g, col_g, val_g, rowptr_o, col_o, val_o, tol=MATRIX_TOL):
    assert np.array_equal(rowptr_g, rowptr_o)
    assert np.array_equal(col_g, col_o)
    n = len(rowptr_o) - 1
    A = sp.csr_matrix((np.abs(val_o), col_o, rowptr_o), shape=(n, n))
    rowmax = np.asarray(A.max(axis=1).todense()).ravel()
    rows = np.repeat(np.arange(n), np.diff(rowptr_o))
    err = np.abs(val_g - val_o) / np.maximum(rowmax[rows], 1e-300)
    assert err.max() <= tol, "max row-relative matrix error %.3e" % err.max()


def nl_state(prob, seed=1):
    L = np.array(prob.mesh.p1) - np.array(prob.mesh.p0)
    u = smooth_field(prob, 0.03 * L.min(), seed)
    du = smooth_field(prob, 0.004 * L.min(), seed + 1)
    v_old = smooth_field(prob, 0.5 * L.min(), seed + 2)
    a_old = smooth_field(prob, 20.0 * L.min(), seed + 3)
    rng = np.random.RandomState(seed)
    traction = rng.uniform(-2e3, 2e3, size=prob.n_iface_nodes * prob.dim)
    return u, du, v_old, a_old, traction


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (2, 1, [3, 5], "cellwise"),
    (2, 2, [3, 5], "cellwise"),
    (2, 2, [3, 4], "component_wise"),
    (3, 1, [2, 3, 2], "cellwise"),
    (3, 2, [2, 3, 2], "cellwise"),
    (3, 2, [2, 2, 3], "component_wise"),
    (3, 2, [3, 2, 2], "lexicographic"),
])
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
    assert rel_err(h.get_vector(capi.NL_ACCELERATION), o.get(orc.NL_ACCELERATION)) < 1e-15
    assert np.array_equal(h.get_vector(capi.NL_EXTERNAL_STRESS), o.get(orc.NL_EXTERNAL_STRESS))
    # bitwise reproducibility (no floating-point atomics): assemble again, compare bits
    v1 = h.export_csr(capi.MAT_TANGENT)[2]
    r1 = h.get_vector(capi.NL_SYSTEM_RHS)
    h.nl_newton_assemble()
    assert np.array_equal(v1, h.export_csr(capi.MAT_TANGENT)[2])
    assert np.array_equal(r1, h.get_vector(capi.NL_SYSTEM_RHS))
    h.close()


def test_chunked_element_buffer_gives_identical_matrix(libs, monkeypatch):
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2)
    prob = make_problem(p, 3, reps=[2, 3, 2])
    u, du, v_old, a_old, traction = nl_state(prob)
    vals = []
    for budget in ("8192", "0.2"):   # 0.2 MB: 3 cells of 3D Q2 per chunk
        monkeypatch.setenv("GF_KE_BUDGET_MB", budget)
        h = capi.Handle(prob)
        h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
        h.set_vector(capi.NL_SOLUTION_DELTA, du)
        h.set_traction(traction)
        h.nl_newton_assemble()
        vals.append(h.export_csr(capi.MAT_TANGENT)[2])
        h.close()
    assert np.array_equal(vals[0], vals[1])


def test_det_F_nonpositive_is_reported(libs):
    capi, solvers, orc = libs
    prob = make_problem(nl_params(poly_degree=1), 2, reps=[2, 2])
    h = capi.Handle(prob)
    x = prob.mesh.support_points
    u = np.zeros(prob.n_dofs)
    u[:] = -2.5 * (x[:, 0] - prob.mesh.p0[0])   # collapses and inverts the mesh in x
    from helpers import dof_components
    u[dof_components(prob) != 0] = 0.0
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    with pytest.raises(capi.GraftError) as e:
        h.nl_newton_assemble()
    assert e.value.code == capi.GF_ERR_DET_F
    h.close()


@pytest.mark.parametrize("dim,reps", [(2, [2, 6]), (3, [2, 3, 2])])
def test_spmv_and_cg_match_oracle_matrix(libs, dim, reps):
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2)
    prob = make_problem(p, dim, reps=reps)
    u, du, v_old, a_old, traction = nl_state(prob)
    h = capi.Handle(prob)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.set_traction(traction)
    h.nl_newton_assemble()
    rowptr, col, val = h.export_csr(capi.MAT_TANGENT)
    A = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs))
    rng = np.random.RandomState(7)
    x = rng.uniform(-1, 1, prob.n_dofs)
    h.set_vector(capi.VEC_SCRATCH0, x)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    y = h.get_vector(capi.VEC_SCRATCH1)
    assert rel_err(y, A @ x) < 1e-13
    # CG (block Jacobi) to a tight tolerance against a sparse direct solve of the same matrix
    b = h.get_vector(capi.NL_SYSTEM_RHS)
    xs = spla.spsolve(A.tocsc(), b)
    for precond in (capi.PRECOND_BLOCK_JACOBI, capi.PRECOND_JACOBI, capi.PRECOND_NONE):
        h.set_option(capi.OPT_PRECONDITIONER, precond)
        h.set_vector(capi.NL_SOLUTION_DELTA, np.zeros(prob.n_dofs))  # same state => same A, b
        h.nl_newton_assemble()
        h.set_vector(capi.NL_NEWTON_UPDATE, np.zeros(prob.n_dofs))
        it, res, upd = h.nl_newton_solve(0, 1e-12, 10.0)
        assert it > 0 and res <= 1e-12 * np.linalg.norm(b)
        assert rel_err(h.get_vector(capi.NL_NEWTON_UPDATE), xs) < 1e-8
    # SolverControl semantics: failure at max iterations is an error, as in the reference
    h.set_vector(capi.NL_SOLUTION_DELTA, np.zeros(prob.n_dofs))
    h.nl_newton_assemble()
    h.set_vector(capi.NL_NEWTON_UPDATE, np.zeros(prob.n_dofs))
    with pytest.raises(capi.GraftError) as e:
        h.nl_newton_solve(0, 1e-14, 3.0 / prob.n_dofs)
    assert e.value.code == capi.GF_ERR_NOT_CONVERGED
    h.close()


def run_nonlinear(libs, prob, n_steps, traction_of_t, n_sub=1):
    capi, solvers, orc = libs
    part = solvers.FakeParticipant(prob.dim, n_steps, prob.params.delta_t, traction_of_t, n_sub)
    solid = solvers.Solid(prob, part)
    solid.run()
    return solid, part


def run_oracle_nonlinear(orc, prob, n_steps, traction_of_t, n_sub=1):
    o = orc.Oracle(prob)
    counts, written = [], []
    dt = prob.params.delta_t
    for w in range(n_steps):
        for it in range(n_sub):
            if n_sub > 1 and it == 0:
                o.save_state()
            o.format_precice_to_deal(traction_of_t((w + 1) * dt, it), orc.NL_EXTERNAL_STRESS)
            n, hist = o.nl_timestep()
            counts.append(n)
            written.append(o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT))
            if n_sub > 1 and it + 1 < n_sub:
                o.reload_state()
    return o, counts, written


@pytest.mark.parametrize("dim,scenario,reps,load", [
    (2, "FSI3", [18, 3], (0.0, -1500.0)),
    (3, "PF", [3, 6, 2], (1500.0, 0.0, 0.0)),
])
def test_newton_counts_and_watchpoint_match_oracle(libs, dim, scenario, reps, load):
    """Identical Newton iteration counts; interface (watch-point) displacement 1e-8 relative.
    Both sides solve the linear systems tightly (the reference's Direct path), so the comparison is
    independent of the preconditioner swap SSOR -> block Jacobi."""
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2, scenario=scenario, type_lin="Direct", delta_t=0.01)
    prob = make_problem(p, dim, reps=reps)
    n = prob.n_iface_nodes

    def traction(t, it):
        ramp = min(1.0, t / 0.03)
        return np.tile(np.array(load) * ramp, n)

    solid, part = run_nonlinear(libs, prob, 4, traction)
    o, counts, written = run_oracle_nonlinear(orc, prob, 4, traction)
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-8
    assert rel_err(solid.handle.get_vector(capi.NL_TOTAL_DISPLACEMENT),
                   o.get(orc.NL_TOTAL_DISPLACEMENT)) < 1e-8
    assert rel_err(solid.handle.get_vector(capi.NL_VELOCITY), o.get(orc.NL_VELOCITY)) < 1e-6
    solid.handle.close()


def test_cg_path_newton_counts_match_oracle_ssor_cg(libs):
    """'Solver type = CG' (tol_lin 1e-6 relative): device block-Jacobi CG vs the oracle's SSOR CG.
    Newton counts must agree; displacements agree to the inexact-Newton level."""
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01)
    prob = make_problem(p, 3, reps=[3, 6, 2])
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([1200.0, 0.0, 0.0]) * min(1.0, t / 0.02), n)
    solid, part = run_nonlinear(libs, prob, 3, traction)
    o, counts, written = run_oracle_nonlinear(orc, prob, 3, traction)
    assert [len(r) for r in solid.history] == counts
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-6
    solid.handle.close()


def test_implicit_coupling_checkpoint_restore(libs):
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2, scenario="PF", type_lin="Direct", delta_t=0.01)
    prob = make_problem(p, 2, reps=[3, 9])
    n = prob.n_iface_nodes
    relax = [0.6, 0.9, 1.0]

    def traction(t, it):
        return np.tile(np.array([900.0, 0.0]) * relax[it], n)

    solid, part = run_nonlinear(libs, prob, 2, traction, n_sub=3)
    o, counts, written = run_oracle_nonlinear(orc, prob, 2, traction, n_sub=3)
    assert [len(r) for r in solid.history] == counts
    assert len(part.written) == 6
    for (w, it, data), ref in zip(part.written, written):
        assert rel_err(data, ref) < 1e-8
    assert solid.time.get_timestep() == 2 and abs(solid.time.current() - 0.02) < 1e-14
    solid.handle.close()


@pytest.mark.parametrize("dim,degree,reps,numbering", [
    (2, 2, [3, 18], "cellwise"),
    (3, 1, [3, 4, 2], "cellwise"),
    (3, 2, [2, 3, 2], "component_wise"),
])
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
    for step in range(3):
        buf = np.tile(load * (step + 1), n)
        o.format_precice_to_deal(buf, orc.LIN_STRESS)
        it_o, res_o = o.lin_step()
        h.set_traction(buf)
        it_g, res_g = h.lin_step(0, p.max_iterations_lin)
        if step == 0:
            assert_matrix_close(*h.export_csr(capi.MAT_SYSTEM), rowptr_o, col_o,
                                o.values(orc.MAT_SYSTEM))
        assert res_g <= 1e-10 and it_g > 0
        assert rel_err(h.get_vector(capi.LIN_OLD_STRESS), o.get(orc.LIN_OLD_STRESS)) < 1e-12
        # both CG solves stop at 1e-10 absolute; velocities agree to solver accuracy
        assert np.abs(h.get_vector(capi.LIN_VELOCITY) - o.get(orc.LIN_VELOCITY)).max() < 1e-7
        d_g, d_o = h.get_vector(capi.LIN_DISPLACEMENT), o.get(orc.LIN_DISPLACEMENT)
        assert np.abs(d_g - d_o).max() < 1e-9
        assert rel_err(h.get_interface_displacement(), o.format_deal_to_precice(orc.LIN_DISPLACEMENT)) < 1e-6
    h.close()


def test_linear_run_with_direct_stand_in_matches_oracle(libs):
    capi, solvers, orc = libs
    p = lin_params(poly_degree=2, type_lin="Direct")
    prob = make_problem(p, 2)   # config 1: PF 2D Q2, 3 x 18 cells, 518 dofs
    n = prob.n_iface_nodes
    traction = lambda t, it: np.tile(np.array([50.0, 0.0]), n)
    part = solvers.FakeParticipant(2, 5, p.delta_t, traction)
    ed = solvers.ElastoDynamics(prob, part)
    ed.run()
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    for step in range(5):
        o.format_precice_to_deal(traction(0, 0), orc.LIN_STRESS)
        o.lin_step()
        assert rel_err(part.written[step][2], o.format_deal_to_precice(orc.LIN_DISPLACEMENT)) < 1e-8
    ed.handle.close()


def test_state_save_restore_and_interface_roundtrip(libs):
    capi, solvers, orc = libs
    prob = make_problem(nl_params(poly_degree=2), 3, reps=[1, 2, 1], numbering="component_wise")
    h = capi.Handle(prob)
    rng = np.random.RandomState(5)
    vecs = {k: rng.uniform(-1, 1, prob.n_dofs) for k in range(6)}
    for k, v in vecs.items():
        h.set_vector(k, v)
    h.state_save()
    for k in range(6):
        h.set_vector(k, np.zeros(prob.n_dofs))
    h.state_restore()
    for k, v in vecs.items():
        assert np.array_equal(h.get_vector(k), v)
    buf = h.get_interface_displacement()
    assert np.array_equal(buf.reshape(-1, 3), vecs[0][prob.iface_dofs.T])
    h.set_traction(buf)
    s = h.get_vector(capi.NL_EXTERNAL_STRESS)
    assert np.array_equal(s[prob.iface_dofs.T], buf.reshape(-1, 3))
    mask = np.ones(prob.n_dofs, bool)
    mask[prob.iface_dofs.reshape(-1)] = False
    assert np.all(s[mask] == 0)
    h2 = capi.Handle(prob)
    with pytest.raises(capi.GraftError):
        h2.state_restore()
    h.close()
    h2.close()


def test_partitioned_assembly_matches_global_rows(libs):
    """slab partitions (owned + one ghost cell layer) assembled independently on one
    GPU reproduce the rows of the single-domain matrix bit for bit (no communicator needed for the
    assembly itself because every rank assembles complete owned rows)."""
    capi, solvers, orc = libs
    p = nl_params(poly_degree=2)
    prob = make_problem(p, 3, reps=[2, 6, 2])
    u, du, v_old, a_old, traction = nl_state(prob)
    h = capi.Handle(prob)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, u)
    h.set_vector(capi.NL_SOLUTION_DELTA, du)
    h.set_traction(traction)
    h.nl_newton_assemble()
    rowptr, col, val = h.export_csr(capi.MAT_TANGENT)
    A = sp.csr_matrix((val, col, rowptr), shape=(prob.n_dofs, prob.n_dofs))
    rhs = h.get_vector(capi.NL_SYSTEM_RHS)
    h.close()
    for nparts in (2, 3):
        seen = np.zeros(prob.n_dofs, dtype=bool)
        for rank in range(nparts):
            part = prob.mesh.partition(1, nparts, rank)
            hp = capi.Handle(prob, partition=part)
            l2g = part.local_to_global
            hp.set_vector(capi.NL_TOTAL_DISPLACEMENT, u[l2g])
            hp.set_vector(capi.NL_SOLUTION_DELTA, du[l2g])
            hp.set_traction(traction.reshape(-1, 3)[hp.iface_visible].reshape(-1))
            hp.nl_newton_assemble()
            rp, cl, vl = hp.export_csr(capi.MAT_TANGENT)
            Ap = sp.csr_matrix((vl, cl, rp), shape=(part.n_owned_dofs, part.n_local_dofs)).tocoo()
            owned_g = l2g[:part.n_owned_dofs]
            assert not seen[owned_g].any()
            seen[owned_g] = True
            Ag = A[owned_g].tocoo()
            Aloc = sp.csr_matrix((Ap.data, (Ap.row, l2g[Ap.col])),
                                 shape=(part.n_owned_dofs, prob.n_dofs))
            diff = (Aloc - A[owned_g]).tocoo()
            assert diff.nnz == 0 or np.abs(diff.data).max() == 0.0
            assert np.array_equal(hp.get_vector(capi.NL_SYSTEM_RHS)[:part.n_owned_dofs], rhs[owned_g])
            hp.close()
        assert seen.all()
