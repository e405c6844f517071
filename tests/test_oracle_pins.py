"""Known-answer tests that pin the CPU oracle independently of deal.II.
The reference has no tests or golden vectors for this path, so these are the substitutes."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import ref_formulas as rf
from helpers import dof_components, lin_params, nl_params, rel_err, smooth_field
from dealii_adapter_b200.problem import make_problem

MU, NU = 0.5e6, 0.4


@pytest.fixture(scope="module")
def orc(native_libs):
    from oracle import oracle_py
    return oracle_py


def random_F(dim, seed, amp=0.2):
    rng = np.random.RandomState(seed)
    return np.eye(dim) + amp * rng.uniform(-1, 1, size=(dim, dim))


@pytest.mark.parametrize("dim", [2, 3])
def test_material_matches_dense_tensor_transcription(orc, dim):
    # compressible_neo_hook_material.h:37-49 vs an independent full-4th-order-tensor transcription
    for seed in range(5):
        F = random_F(dim, seed)
        J = np.linalg.det(F)
        Fb = J ** (-1.0 / dim) * F
        b_bar = rf.to_voigt2(Fb @ Fb.T)
        psi, tau, Jc = orc.material(dim, MU, NU, J, b_bar)
        tau_ref, Jc_ref = rf.tau_Jc(F, MU, NU)
        assert abs(psi - rf.psi(F, MU, NU)) <= 1e-12 * abs(psi)
        assert rel_err(tau, rf.to_voigt2(tau_ref)) < 1e-13
        assert rel_err(Jc, rf.to_voigt4(Jc_ref)) < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_tau_is_derivative_of_reference_psi(orc, dim):
    # tau = dPsi/dF F^T (Kirchhoff stress of the reference's own strain energy, material.h:62-72)
    F = random_F(dim, 11)
    eps = 1e-6
    dPsi = np.zeros((dim, dim))
    for i in range(dim):
        for j in range(dim):
            Fp, Fm = F.copy(), F.copy()
            Fp[i, j] += eps
            Fm[i, j] -= eps
            dPsi[i, j] = (rf.psi(Fp, MU, NU) - rf.psi(Fm, MU, NU)) / (2 * eps)
    tau_fd = dPsi @ F.T
    J = np.linalg.det(F)
    Fb = J ** (-1.0 / dim) * F
    _, tau, _ = orc.material(dim, MU, NU, J, rf.to_voigt2(Fb @ Fb.T))
    assert rel_err(tau, rf.to_voigt2(0.5 * (tau_fd + tau_fd.T))) < 1e-7


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2), (2, 3), (2, 4), (3, 3)])
def test_cell_residual_and_tangent_are_derivatives_of_element_energy(orc, dim, degree):
    # nonlinear_elasticity.cc:984-985 (internal force) and :1011-1023 (tangent) against finite
    # differences of Pi(u) = sum_q Psi(F_q) JxW built from the reference's Psi. rho = 0 removes the
    # inertia terms; body force = 0.
    p = nl_params(poly_degree=degree, rho=0.0, scenario="PF")
    prob = make_problem(p, dim, reps=[1] * dim)
    o = orc.Oracle(prob, n_threads=1)
    dpc = prob.mesh.dofs_per_cell
    h = (np.array(prob.mesh.p1) - np.array(prob.mesh.p0))
    rng = np.random.RandomState(3)
    u = 0.02 * h.min() * rng.uniform(-1, 1, size=dpc)
    K, r = o.nl_cell(0, u, np.zeros(dpc))
    nq1 = degree + 2
    E = lambda v: rf.element_energy(v, h, dim, degree, nq1, MU, NU)
    eps = 1e-6 * h.min()
    g = np.zeros(dpc)
    for i in range(dpc):
        e = np.zeros(dpc)
        e[i] = eps
        g[i] = (E(u + e) - E(u - e)) / (2 * eps)
    assert rel_err(-r, g) < 2e-6
    assert np.abs(K - K.T).max() <= 1e-12 * np.abs(K).max()
    # tangent: FD of the oracle's own residual (2nd differences of Pi are too noisy)
    Kfd = np.zeros((dpc, dpc))
    for j in range(dpc):
        e = np.zeros(dpc)
        e[j] = eps
        _, rp = o.nl_cell(0, u + e, np.zeros(dpc))
        _, rm = o.nl_cell(0, u - e, np.zeros(dpc))
        Kfd[:, j] = -(rp - rm) / (2 * eps)
    assert rel_err(K, Kfd) < 1e-6


@pytest.mark.parametrize("dim,degree", [(2, 2), (3, 2), (2, 3), (2, 4), (3, 3)])
def test_rigid_motion_gives_zero_internal_force(orc, dim, degree):
    p = nl_params(poly_degree=degree, rho=0.0)
    prob = make_problem(p, dim, reps=[1] * dim)
    o = orc.Oracle(prob, n_threads=1)
    dpc = prob.mesh.dofs_per_cell
    nodes = rf.support_points_1d(degree)[np.array(rf.hierarchical_nodes(dim, degree))]
    h = np.array(prob.mesh.p1) - np.array(prob.mesh.p0)
    X = nodes * h
    th = 0.3
    R = np.eye(dim)
    R[0, 0], R[0, 1], R[1, 0], R[1, 1] = np.cos(th), -np.sin(th), np.sin(th), np.cos(th)
    un = X @ R.T - X + 0.01                                  # [node, component]
    u = np.array([un[a, c] for a, c in rf.system_to_node_component(dim, degree)])
    K, r = o.nl_cell(0, u, np.zeros(dpc))
    assert np.abs(r).max() < 1e-9 * MU * h.max() ** (dim - 1)


@pytest.mark.parametrize("degree,reps", [(2, [2, 3, 2]), (3, [1, 2, 1])])
def test_nonlinear_tangent_at_zero_equals_linear_stiffness_plus_mass_3d(orc, degree, reps):
    # K_nl(u=0) - alpha_1 M == K_lin in 3D (lambda_eff = kappa - 2mu/3 = lambda)
    pn = nl_params(poly_degree=degree)
    pl = lin_params(poly_degree=degree, delta_t=pn.delta_t)
    probn = make_problem(pn, 3, reps=reps)
    probl = make_problem(pl, 3, reps=reps)
    # no constraints so that both matrices are the raw sums
    probn.constrained[:] = 0
    probl.constrained[:] = 0
    on, ol = orc.Oracle(probn), orc.Oracle(probl)
    on.nl_assemble_system()
    ol.lin_assemble_system()
    alpha_1 = 1.0 / (pn.beta * pn.delta_t ** 2)
    Knl = on.values(orc.MAT_TANGENT)
    K = ol.values(orc.MAT_STIFFNESS)
    M = ol.values(orc.MAT_MASS)
    assert rel_err(Knl, K + alpha_1 * M) < 1e-12
    # mass matrix row sums = rho * |support| partition of unity: total = rho*|Omega| per component
    vol = np.prod(np.array(probl.mesh.p1) - np.array(probl.mesh.p0))
    assert abs(M.sum() - 3 * pl.rho * vol) < 1e-10 * pl.rho * vol


@pytest.mark.parametrize("degree", [2, 3, 4])
def test_neumann_load_integrates_traction_over_interface(orc, degree):
    # nonlinear_elasticity.cc:791-859 at u=0 (pull-back factor = 1): sum of nodal loads per
    # component = traction * interface area ; same for the linear consistent loading :458-521
    for dim in ((2, 3) if degree < 4 else (2,)):
        pn = nl_params(poly_degree=degree)
        prob = make_problem(pn, dim, reps=[2, 3, 2][:dim])
        prob.constrained[:] = 0
        o = orc.Oracle(prob)
        t = np.array([3.0, -2.0, 0.5])[:dim]
        buf = np.tile(t, prob.n_iface_nodes)
        o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
        o.nl_assemble_system()
        rhs = o.get(orc.NL_SYSTEM_RHS)
        L = np.array(prob.mesh.p1) - np.array(prob.mesh.p0)
        if dim == 2:
            area = 2 * L[1] + L[0]          # x-, x+, y+ of the PF flap
        else:
            area = 2 * L[1] * L[2] + L[0] * L[2]
        comp = dof_components(prob)
        for c in range(dim):
            assert abs(rhs[comp == c].sum() - t[c] * area) < 1e-12 * abs(t).max() * area
        pl = lin_params(poly_degree=degree)
        probl = make_problem(pl, dim, reps=[2, 3, 2][:dim])
        ol = orc.Oracle(probl)
        ol.format_precice_to_deal(buf, orc.LIN_STRESS)
        ol.lin_assemble_system()
        ol.lin_assemble_rhs()
        # F_{n+1} is kept in old_stress (linear_elasticity.cc:405,409)
        F = ol.get(orc.LIN_OLD_STRESS)
        compl = dof_components(probl)
        for c in range(dim):
            assert abs(F[compl == c].sum() - t[c] * area) < 1e-12 * abs(t).max() * area


def test_ssor_cg_solves_the_system(orc):
    # SolverCG + precondition_SSOR restatement against a sparse direct solve
    p = nl_params(poly_degree=2, tol_lin=1e-12)
    prob = make_problem(p, 2, reps=[2, 6])
    o = orc.Oracle(prob)
    buf = np.tile([1000.0, 0.0], prob.n_iface_nodes)
    o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
    o.nl_update_acceleration()
    o.nl_assemble_system()
    st, it, res = o.nl_solve_linear_system()
    assert st == 0 and it > 0
    A = o.csr(orc.MAT_TANGENT).tocsc()
    b = o.get(orc.NL_SYSTEM_RHS)
    x = spla.spsolve(A, b)
    assert rel_err(o.get(orc.NL_NEWTON_UPDATE), x) < 1e-9
    # constrained rows: diagonal only, zero rhs (distribute_local_to_global semantics)
    c = np.flatnonzero(prob.constrained)
    Ad = A.toarray()
    assert np.all(b[c] == 0)
    off = Ad[c].copy()
    off[np.arange(len(c)), c] = 0
    assert np.all(off == 0) and np.all(Ad[:, c][np.setdiff1d(np.arange(prob.n_dofs), c)] == 0)
    assert np.all(Ad[c, c] > 0)


def test_linear_rhs_matches_theta_scheme_formula(orc):
    # linear_elasticity.cc:397-420 against the closed form with the oracle's own K, M
    p = lin_params(poly_degree=2)
    prob = make_problem(p, 2)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    K = o.csr(orc.MAT_STIFFNESS)
    M = o.csr(orc.MAT_MASS)
    rng = np.random.RandomState(0)
    v = rng.uniform(-1, 1, prob.n_dofs) * (prob.constrained == 0)
    d = 1e-3 * rng.uniform(-1, 1, prob.n_dofs) * (prob.constrained == 0)
    Fold = rng.uniform(-1, 1, prob.n_dofs)
    o.set(orc.LIN_VELOCITY, v)
    o.set(orc.LIN_DISPLACEMENT, d)
    o.set(orc.LIN_OLD_STRESS, Fold)
    buf = np.tile([100.0, -50.0], prob.n_iface_nodes)
    o.format_precice_to_deal(buf, orc.LIN_STRESS)
    o.lin_assemble_rhs()
    Fnew = o.get(orc.LIN_OLD_STRESS)
    dt, th = p.delta_t, p.theta
    rhs = dt * th * Fnew + dt * (1 - th) * Fold + M @ v - th * (1 - th) * dt * dt * (K @ v) - dt * (K @ d)
    rhs[prob.constrained != 0] = 0
    assert rel_err(o.get(orc.LIN_SYSTEM_RHS), rhs) < 1e-12
    A = o.csr(orc.MAT_SYSTEM)
    S = (M + th * th * dt * dt * K).tolil()
    c = np.flatnonzero(prob.constrained)
    diag = S.diagonal()
    S[c, :] = 0
    S[:, c] = 0
    S[c, c] = diag[c]
    assert abs(A - S.tocsr()).max() < 1e-12 * abs(A).max()
    st, it, res = o.lin_solve()
    assert st == 0 and res <= 1e-10
    x = spla.spsolve(A.tocsc(), o.get(orc.LIN_SYSTEM_RHS))
    assert rel_err(o.get(orc.LIN_VELOCITY), x) < 1e-8


def test_checkpoint_and_interface_format_roundtrip(orc):
    p = nl_params(poly_degree=2)
    prob = make_problem(p, 3, reps=[1, 2, 1])
    o = orc.Oracle(prob)
    rng = np.random.RandomState(5)
    u = rng.uniform(-1, 1, prob.n_dofs)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    buf = o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT)
    assert np.array_equal(buf.reshape(-1, 3), u[prob.iface_dofs.T])
    o.save_state()
    o.set(orc.NL_TOTAL_DISPLACEMENT, np.zeros(prob.n_dofs))
    o.reload_state()
    assert np.array_equal(o.get(orc.NL_TOTAL_DISPLACEMENT), u)
    o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
    s = o.get(orc.NL_EXTERNAL_STRESS)
    assert np.array_equal(s[prob.iface_dofs.T], buf.reshape(-1, 3))
    mask = np.ones(prob.n_dofs, bool)
    mask[prob.iface_dofs.reshape(-1)] = False
    assert np.all(s[mask] == 0)
