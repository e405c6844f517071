"""Golden vectors produced by THE REFERENCE'S OWN CODE, run in this container.

The pieces of /root/reference that carry the arithmetic and the control flow of the hot path
compile without the real deal.II against the small stand-in oracle/ref_shim (`make -C oracle ref`
-> oracle/_ref/*, sources taken in place; blocks that live inside the big translation units are cut
out by marker at build time and never committed):
  ref_driver             compressible_neo_hook_material.h (Psi, tau, Jc), postprocessor.h, time_handler.h
  ref_assembler_driver   Assembler_Base / Assembler<dim,double> + PointHistory: K_e, r_e, Neumann term
  ref_linear_driver      local stiffness loops and consistent-loading face loop of linear_elasticity.cc
  ref_updates_driver     Newmark coefficients / updates / masked norms; theta-scheme rhs + update
  ref_newton_driver      Solid::solve_nonlinear_timestep + Errors with scripted norms
  ref_run_driver         Solid::run() / ElastoDynamics::run() against recording stand-ins
  ref_adapter_driver     Adapter::format_* and checkpoint members
  ref_grid_driver        make_grid of both solvers against a recording grid generator
  ref_constraints_driver make_constraints / boundary-value block against a recording VectorTools
  ref_solver_driver      solve_linear_system / solve against recording SolverControl / SolverCG
  ref_parameters_driver  include/adapter/parameters.{h,cc} unmodified, against a ParameterHandler stand-in
This script runs them on deterministic inputs and writes tests/golden/reference_vectors.npz;
tests/test_reference_pins.py, tests/test_host_driver_cpu.py and tests/test_gpu_zz_reference_pins.py
check the oracle, the host mirrors and the device against the file, and — where /root/reference
exists — that the file is what the reference produces.

usage: python tests/golden/make_reference_vectors.py [--check]
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
ASM_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_assembler_driver")
LIN_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_linear_driver")
UPD_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_updates_driver")
NEWTON_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_newton_driver")
ADAPTER_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_adapter_driver")
GRID_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_grid_driver")
PRM_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_parameters_driver")
RUN_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_run_driver")
CONSTRAINTS_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_constraints_driver")
SOLVER_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_solver_driver")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "reference_vectors.npz")


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                          stdout=subprocess.DEVNULL)


def run(args, stdin=None):
    return subprocess.run([DRIVER] + [repr(float(a)) if isinstance(a, float) else str(a) for a in args],
                          input=stdin, capture_output=True, text=True, check=True).stdout


def material_cases():
    rng = np.random.RandomState(20260117)
    cases = []
    for dim in (2, 3):
        n = dim * (dim + 1) // 2
        for mu, nu in ((0.5e6, 0.4), (1538462.0, 0.3), (2.0e6, 0.0), (1.0e5, 0.49)):
            for _ in range(6):
                # b_bar = F_bar F_bar^T of a random deformation gradient (SPD, det 1), det F apart
                F = np.eye(dim) + 0.35 * rng.uniform(-1, 1, (dim, dim))
                if np.linalg.det(F) <= 0.2:
                    F = np.eye(dim) + 0.1 * rng.uniform(-1, 1, (dim, dim))
                J = np.linalg.det(F)
                Fb = J ** (-1.0 / dim) * F
                b = Fb @ Fb.T
                idx = [(0, 0), (1, 1), (0, 1)] if dim == 2 else \
                    [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
                bv = [float(b[i, j]) for i, j in idx]
                cases.append((dim, mu, nu, float(J), bv))
    return cases


def face_tables(dim, p, nq1, h, faces):
    """Per face: shape values at the face quadrature points, JxW and the outward unit normal of a
    Cartesian cell with edge lengths h. Face points in QProjector order (standard orientation;
    3D y-faces run (z, x)), restated from deal.II as in the oracle."""
    import ref_formulas as rf
    x1, w1 = rf.gauss01(nq1)
    nodes = rf.hierarchical_nodes(dim, p)
    nqf = nq1 ** (dim - 1)
    out = {}
    for f in faces:
        d, c = f // 2, float(f % 2)
        Nf = np.zeros((nqf, len(nodes)))
        JxWf = np.zeros(nqf)
        for q in range(nqf):
            fi = [(q // nq1 ** k) % nq1 for k in range(dim - 1)]
            fq = [x1[i] for i in fi]
            w = np.prod([w1[i] for i in fi])
            xi = [0.0] * dim
            if dim == 2:
                xi[d], xi[1 - d] = c, fq[0]
            elif d == 0:
                xi = [c, fq[0], fq[1]]
            elif d == 1:
                xi = [fq[1], c, fq[0]]
            else:
                xi = [fq[0], fq[1], c]
            v1 = [rf.lagrange_1d(p, xi[k])[0][0] for k in range(dim)]
            for a, lex in enumerate(nodes):
                Nf[q, a] = np.prod([v1[k][lex[k]] for k in range(dim)])
            JxWf[q] = w * np.prod([h[k] for k in range(dim) if k != d])
        normal = np.zeros((nqf, dim))
        normal[:, d] = 1.0 if f % 2 else -1.0
        out[f] = (Nf, JxWf, normal)
    return out


def q1_vertex_basis(dim, xi):
    """values and unit-cell gradients [nv], [nv, dim] of the MappingQ1 vertex shape functions"""
    nv = 1 << dim
    val, grad = np.ones(nv), np.ones((nv, dim))
    for v in range(nv):
        for l in range(dim):
            hi = (v >> l) & 1
            val[v] *= xi[l] if hi else 1.0 - xi[l]
            for k in range(dim):
                grad[v, k] *= (1.0 if hi else -1.0) if l == k else (xi[l] if hi else 1.0 - xi[l])
    return val, grad


def distorted_assembly_cases():
    """One GENERAL (non-affine) cell each: (dim, degree, vertices [nv, dim], interface faces, body
    force). The reference block is geometry-agnostic (it reads FEValues); the tables it gets here
    are those of MappingQ1 on the distorted cell, computed with numpy."""
    rng = np.random.RandomState(77)
    out = []
    for dim, p, h, bf in ((2, 2, [0.03, 0.02], (0.0, -9.81, 0.0)), (3, 1, [0.05, 0.04, 0.06], (1.5, -9.81, 0.5)),
                          (2, 3, [0.04, 0.05], (0.0, 0.0, 0.0))):
        nv = 1 << dim
        verts = np.array([[((v >> d) & 1) * h[d] for d in range(dim)] for v in range(nv)], dtype=float)
        verts += 0.15 * np.array(h) * rng.uniform(-1, 1, size=(nv, dim))
        out.append((dim, p, verts, [0, 1, 3], bf))
    return out


def run_distorted_assembly_case(k, dim, p, verts, faces, body_force):
    import ref_formulas as rf
    rng = np.random.RandomState(300 + k)
    nq1 = p + 2
    N, dN, w = rf.cell_tables(dim, p, nq1)
    x1, w1 = rf.gauss01(nq1)
    npc, nq, nqf = N.shape[1], len(w), nq1 ** (dim - 1)
    dpc = npc * dim
    gradN, JxW = np.zeros_like(dN), np.zeros(nq)
    for q in range(nq):
        xi = [x1[(q // nq1 ** d) % nq1] for d in range(dim)]
        _, g = q1_vertex_basis(dim, xi)
        J = verts.T @ g                                     # J[i, j] = dX_i / dxi_j
        gradN[q] = dN[q] @ np.linalg.inv(J)
        JxW[q] = np.linalg.det(J) * w[q]
    assert JxW.min() > 0
    nodes = rf.hierarchical_nodes(dim, p)
    mu, nu, rho, beta, dt = 0.5e6, 0.4, 1000.0, 0.25, 0.01
    alpha_1 = 1.0 / (beta * dt * dt)
    hmin = np.abs(verts[-1] - verts[0]).min()
    u = (0.1 if p <= 2 else 0.1 / (p * p)) * hmin * rng.uniform(-1, 1, dpc)
    acc = 50.0 * rng.uniform(-1, 1, dpc)
    stress = 2000.0 * rng.uniform(-1, 1, dpc)
    words = [dim, npc, nq, nqf, len(faces), mu, nu, rho, alpha_1] + list(body_force) + [7]
    words += list(N.reshape(-1)) + list(gradN.reshape(-1)) + list(JxW)
    for f in faces:
        d, c = f // 2, float(f % 2)
        Nf, JxWf, normal = np.zeros((nqf, npc)), np.zeros(nqf), np.zeros((nqf, dim))
        for q in range(nqf):
            fi = [(q // nq1 ** kk) % nq1 for kk in range(dim - 1)]
            fq = [x1[i] for i in fi]
            wq = np.prod([w1[i] for i in fi])
            if dim == 2:
                xi = [0.0, 0.0]
                xi[d], xi[1 - d] = c, fq[0]
            elif d == 0:
                xi = [c, fq[0], fq[1]]
            elif d == 1:
                xi = [fq[1], c, fq[0]]
            else:
                xi = [fq[0], fq[1], c]
            v1 = [rf.lagrange_1d(p, xi[kk])[0][0] for kk in range(dim)]
            for a, lex in enumerate(nodes):
                Nf[q, a] = np.prod([v1[kk][lex[kk]] for kk in range(dim)])
            _, g = q1_vertex_basis(dim, xi)
            J = verts.T @ g
            n_ref = np.zeros(dim)
            n_ref[d] = 1.0 if f % 2 else -1.0
            nda = np.linalg.det(J) * np.linalg.inv(J).T @ n_ref      # Nanson: n da
            JxWf[q] = np.linalg.norm(nda) * wq
            normal[q] = nda / np.linalg.norm(nda)
        words += [f, 7] + list(Nf.reshape(-1)) + list(JxWf) + list(normal.reshape(-1))
    words += list(u) + list(acc) + list(stress)
    if p > 2:
        s2c = rf.system_to_node_component(dim, p)
        words += [a for a, c in s2c] + [c for a, c in s2c]
    text = " ".join(repr(float(x)) if isinstance(x, (float, np.floating)) else str(int(x)) for x in words)
    res = subprocess.run([ASM_DRIVER], input=text, capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in res.strip().split("\n")]
    K = np.array(rows[:dpc], dtype=float)
    r = np.array(rows[dpc], dtype=float)
    meta = np.array([dim, p, 0.0, 0.0, 0.0] + list(body_force) + [mu, nu, rho, beta, dt])
    return {"asm%d_meta" % k: meta, "asm%d_faces" % k: np.array(faces, dtype=np.int64),
            "asm%d_verts" % k: verts, "asm%d_u" % k: u, "asm%d_acc" % k: acc,
            "asm%d_stress" % k: stress, "asm%d_K" % k: K, "asm%d_r" % k: r}


def assembly_cases():
    """One Cartesian cell each: (dim, degree, edge lengths, interface faces, body force)."""
    return [(2, 1, [0.1, 0.05], [0, 1, 3], (0.0, 0.0, 0.0)),
            (2, 2, [0.02, 0.03], [0, 1, 3], (0.0, -9.81, 0.0)),
            (3, 1, [0.1, 0.2, 0.15], [0, 1, 3], (0.0, 0.0, 0.0)),
            (3, 2, [0.05, 0.04, 0.06], [0, 1, 3], (1.5, -9.81, 0.5)),
            (3, 2, [0.004, 0.007, 0.0125], [], (0.0, 0.0, 0.0)),
            # degrees of the shipped parameter files (parameters.prm:21, nonlinear_elasticity.prm:24)
            (2, 3, [0.03, 0.02], [0, 1, 3], (0.0, -9.81, 0.0)),
            (2, 4, [0.05, 0.04], [0, 1, 3], (0.0, 0.0, 0.0)),
            (3, 3, [0.05, 0.04, 0.06], [0, 1, 3], (1.5, -9.81, 0.5))]


def run_assembly_case(k, dim, p, h, faces, body_force):
    """Feeds the reference's assemble_system_one_cell (cut out of nonlinear_elasticity.cc at build
    time) with numpy FE tables of tests/ref_formulas.py; returns inputs and its K_e, r_e."""
    import ref_formulas as rf
    rng = np.random.RandomState(100 + k)
    nq1 = p + 2                                      # QGauss(degree + 2), nonlinear_elasticity.cc:74-75
    N, dN, w = rf.cell_tables(dim, p, nq1)
    npc, nq, nqf = N.shape[1], len(w), nq1 ** (dim - 1)
    dpc = npc * dim
    gradN = dN / np.asarray(h)[None, None, :]
    JxW = w * np.prod(h)
    mu, nu, rho, beta, dt = 0.5e6, 0.4, 1000.0, 0.25, 0.01
    alpha_1 = 1.0 / (beta * dt * dt)
    # random nodal values: smaller for the closely spaced Gauss-Lobatto nodes (det F must stay > 0)
    u = (0.15 if p <= 2 else 0.15 / (p * p)) * min(h) * rng.uniform(-1, 1, dpc)
    acc = 50.0 * rng.uniform(-1, 1, dpc)
    stress = 2000.0 * rng.uniform(-1, 1, dpc)
    ft = face_tables(dim, p, nq1, h, faces)
    words = [dim, npc, nq, nqf, len(faces), mu, nu, rho, alpha_1] + list(body_force) + [7]
    words += list(N.reshape(-1)) + list(gradN.reshape(-1)) + list(JxW)
    for f in faces:
        Nf, JxWf, normal = ft[f]
        words += [f, 7] + list(Nf.reshape(-1)) + list(JxWf) + list(normal.reshape(-1))
    words += list(u) + list(acc) + list(stress)
    if p > 2:       # FESystem local numbering beyond node-major (system_to_component_index)
        s2c = rf.system_to_node_component(dim, p)
        words += [a for a, c in s2c] + [c for a, c in s2c]
    text = " ".join(repr(float(x)) if isinstance(x, (float, np.floating)) else str(int(x)) for x in words)
    res = subprocess.run([ASM_DRIVER], input=text, capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in res.strip().split("\n")]
    K = np.array(rows[:dpc], dtype=float)
    r = np.array(rows[dpc], dtype=float)
    meta = np.array([dim, p] + list(h) + [0.0] * (3 - dim) + list(body_force) + [mu, nu, rho, beta, dt])
    return {"asm%d_meta" % k: meta, "asm%d_faces" % k: np.array(faces, dtype=np.int64),
            "asm%d_u" % k: u, "asm%d_acc" % k: acc, "asm%d_stress" % k: stress,
            "asm%d_K" % k: K, "asm%d_r" % k: r}


def run_linear_case(k, dim, p, h, faces):
    """The reference's local stiffness loops (linear_elasticity.cc:289-323) and the face loop of
    assemble_consistent_loading (:487-512) on one cell; QGauss(degree + 1) (:61,252,465)."""
    import ref_formulas as rf
    rng = np.random.RandomState(200 + k)
    nq1 = p + 1
    N, dN, w = rf.cell_tables(dim, p, nq1)
    npc, nq, nqf = N.shape[1], len(w), nq1 ** (dim - 1)
    dpc = npc * dim
    gradN = dN / np.asarray(h)[None, None, :]
    JxW = w * np.prod(h)
    mu, nu = 0.5e6, 0.4
    lam = 2 * mu * nu / (1 - 2 * nu)                 # parameters.cc:189
    stress = 2000.0 * rng.uniform(-1, 1, dpc)
    ft = face_tables(dim, p, nq1, h, faces)
    words = [dim, npc, nq, nqf, len(faces), lam, mu, 6]
    words += list(gradN.reshape(-1)) + list(JxW)
    for f in faces:
        Nf, JxWf, normal = ft[f]
        words += [f, 6] + list(Nf.reshape(-1)) + list(JxWf)
    words += list(stress)
    if p > 2:
        s2c = rf.system_to_node_component(dim, p)
        words += [a for a, c in s2c] + [c for a, c in s2c]
    text = " ".join(repr(float(x)) if isinstance(x, (float, np.floating)) else str(int(x)) for x in words)
    res = subprocess.run([LIN_DRIVER], input=text, capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in res.strip().split("\n")]
    meta = np.array([dim, p] + list(h) + [0.0] * (3 - dim) + [mu, nu])
    return {"lin%d_meta" % k: meta, "lin%d_faces" % k: np.array(faces, dtype=np.int64),
            "lin%d_stress" % k: stress, "lin%d_K" % k: np.array(rows[:dpc], dtype=float),
            "lin%d_F" % k: np.array(rows[dpc], dtype=float)}


def linear_cases():
    return [(2, 1, [0.1, 0.05], [0, 1, 3]), (2, 2, [0.1 / 3, 1.0 / 18], [0, 1, 3]),
            (3, 1, [0.1, 0.2, 0.15], [0, 1, 3]), (3, 2, [0.05, 0.04, 0.06], [0, 1, 3]),
            (2, 3, [0.1 / 3, 1.0 / 18], [0, 1, 3]), (2, 4, [0.05, 0.04], [0, 1, 3]),
            (3, 3, [0.05, 0.04, 0.06], [0, 1, 3])]


def fmt(words):
    return " ".join(w if isinstance(w, str) else
                    (repr(float(w)) if isinstance(w, (float, np.floating)) else str(int(w))) for w in words)


def run_update_cases():
    """The reference's Newmark coefficients / updates / masked norms (nonlinear_elasticity.h:242-250,
    .cc:549-622) and its theta-scheme right-hand side + displacement update (linear_elasticity.cc:
    384-420, 579-586) on small vectors. The constraint mask and the K, M matrices of the linear case
    are taken from a small structured problem (stored in the fixture)."""
    from helpers import lin_params, nl_params
    from dealii_adapter_b200.problem import make_problem
    from oracle import oracle_py as orc
    out = {}
    for k, (beta, gamma, dt) in enumerate(((0.25, 0.5, 0.01), (0.3025, 0.6, 0.005))):
        prob = make_problem(nl_params(poly_degree=1, beta=beta, gamma=gamma, delta_t=dt), 2, reps=[3, 2])
        n = prob.n_dofs
        rng = np.random.RandomState(300 + k)
        vecs = [rng.uniform(-1, 1, n) * s for s in (1e-3, 1e-2, 0.5, 20.0, 1e3, 1e-4)]
        words = ["nl", n, beta, gamma, dt] + [int(c) for c in prob.constrained]
        for v in vecs:
            words += list(v)
        res = subprocess.run([UPD_DRIVER], input=fmt(words), capture_output=True, text=True,
                             check=True).stdout.strip().split("\n")
        out["upd%d_in" % k] = np.array([beta, gamma, dt])
        out["upd%d_vecs" % k] = np.array(vecs)
        out["upd%d_alpha" % k] = np.array(res[0].split(), dtype=float)
        out["upd%d_acc" % k] = np.array(res[1].split(), dtype=float)
        out["upd%d_vel" % k] = np.array(res[2].split(), dtype=float)
        out["upd%d_total" % k] = np.array(res[3].split(), dtype=float)
        out["upd%d_norms" % k] = np.array(res[4].split(), dtype=float)
        assert res[5].strip() == "1"
    for k, (consistent, bf) in enumerate(((False, (0.0, -9.81, 0.0)), (True, (0.0, 0.0, 0.0)))):
        p = lin_params(poly_degree=1, theta=0.6 if k == 0 else 0.5, delta_t=0.005,
                       read_data_name="Stress" if consistent else "Force", body_force=bf)
        prob = make_problem(p, 2, reps=[3, 2])
        n = prob.n_dofs
        o = orc.Oracle(prob, n_threads=1)
        o.lin_assemble_system()
        K = o.csr(orc.MAT_STIFFNESS).toarray()
        M = o.csr(orc.MAT_MASS).toarray()
        rng = np.random.RandomState(400 + k)
        stress, old_stress, vel, disp, new_vel = (rng.uniform(-1, 1, n) * s
                                                  for s in (500.0, 400.0, 0.3, 1e-3, 0.25))
        if consistent:      # what assemble_consistent_loading() leaves in system_rhs for `stress`
            o2 = orc.Oracle(prob, n_threads=1)
            o2.lin_assemble_system()
            o2.set(orc.LIN_STRESS, stress)
            o2.lin_assemble_rhs()
            loading = o2.get(orc.LIN_OLD_STRESS)
        else:
            loading = np.zeros(n)
        bfv = o.get(orc.LIN_BODY_FORCE)
        words = ["lin", n, p.theta, p.delta_t, int(consistent), int(any(bf))]
        words += list(K.reshape(-1)) + list(M.reshape(-1))
        for v in (loading, stress, old_stress, vel, disp, bfv, new_vel):
            words += list(v)
        res = subprocess.run([UPD_DRIVER], input=fmt(words), capture_output=True, text=True,
                             check=True).stdout.strip().split("\n")
        out["rhs%d_in" % k] = np.array([p.theta, p.delta_t, float(consistent)] + list(bf))
        out["rhs%d_K" % k], out["rhs%d_M" % k] = K, M
        out["rhs%d_vecs" % k] = np.array([loading, stress, old_stress, vel, disp, bfv, new_vel])
        for name, line in zip(("system_rhs", "old_stress", "old_velocity", "old_displacement",
                               "displacement"), res):
            out["rhs%d_%s" % (k, name)] = np.array(line.split(), dtype=float)
    return out


def newton_scripts():
    """(max_iterations_NR, tol_f, tol_u, residual norms per assembly, update norms per solve)."""
    cases = [
        (10, 1e-9, 1e-6, [1.0, 0.1, 1e-4, 1e-11, 1e-12], [0.5, 1e-3, 1e-8, 1e-13]),
        (10, 1e-9, 1e-6, [0.0] * 4, [0.0] * 4),                       # nothing to do: norms of 0
        (10, 1e-9, 1e-6, [1e-9, 4e-9, 4e-9], [1e-16, 5e-16, 5e-16]),  # absolute criteria only
        (5, 1e-9, 1e-6, [1.0] * 6, [1.0] * 6),                        # never converges
        (10, 1e-9, 1e-6, [3.0, 2.0, 1e-10, 1e-10, 1e-12], [1.0, 1e-9, 1e-3, 1e-9, 1e-9]),
        (1, 1e-9, 1e-6, [1.0, 1e-12], [1e-12]),                       # one iteration allowed
        (3, 1e-9, 1e-6, [1.0, 1e-3, 1e-12, 1e-13], [1.0, 1e-4, 1e-9]),  # converges at the last check
        (10, 1e-6, 1e-3, [5.0, 4.9e-6, 1e-7], [2.0, 1.9e-3, 1e-7]),   # just inside the tolerances
        (10, 1e-6, 1e-3, [5.0, 5.1e-6, 1e-7, 1e-8], [2.0, 2.1e-3, 1e-7, 1e-9]),  # just outside
    ]
    rng = np.random.RandomState(55)
    for _ in range(8):
        n = 12
        res = np.abs(rng.lognormal(0, 1)) * 10.0 ** (-rng.uniform(0.5, 4.0, n).cumsum())
        upd = np.abs(rng.lognormal(0, 1)) * 10.0 ** (-rng.uniform(0.5, 4.0, n).cumsum())
        cases.append((10, 1e-9, 1e-6, [float(x) for x in res], [float(x) for x in upd]))
    return cases


def run_newton_scripts():
    out = {}
    cases = newton_scripts()
    for k, (max_it, tol_f, tol_u, res, upd) in enumerate(cases):
        n = max(len(res), len(upd), max_it + 1)
        res = list(res) + [res[-1]] * (n - len(res))
        upd = list(upd) + [upd[-1]] * (n - len(upd))
        txt = subprocess.run([NEWTON_DRIVER], input=fmt([max_it, tol_f, tol_u, n] + res + upd),
                             capture_output=True, text=True, check=True).stdout.split("\n")
        solves, assemblies, converged = (int(x) for x in txt[0].split())
        out["newton%02d_in" % k] = np.array([max_it, tol_f, tol_u])
        out["newton%02d_res" % k] = np.array(res)
        out["newton%02d_upd" % k] = np.array(upd)
        out["newton%02d_out" % k] = np.array([solves, assemblies, converged])
        out["newton%02d_last" % k] = np.array(txt[1].split(), dtype=float)
    out["n_newton"] = np.array(len(cases))
    return out


def run_adapter_cases():
    """The reference's Adapter::format_deal_to_precice / format_precice_to_deal / checkpoint
    members (adapter.h:389-489) on the interface index sets of two small problems, with
    component_wise numbering (the reference renumbers component-wise, nonlinear:318)."""
    from helpers import nl_params
    from dealii_adapter_b200.problem import make_problem
    out = {}
    for k, (dim, reps, degree) in enumerate(((2, [3, 4], 2), (3, [2, 3, 2], 1))):
        prob = make_problem(nl_params(poly_degree=degree), dim, reps=reps, numbering="component_wise")
        n, ni = prob.n_dofs, prob.n_iface_nodes
        rng = np.random.RandomState(500 + k)
        vec = rng.uniform(-1, 1, n)
        buf = rng.uniform(-1, 1, dim * ni)
        words = [dim, n, ni] + [int(i) for i in prob.iface_dofs.reshape(-1)] + list(vec) + list(buf)
        res = subprocess.run([ADAPTER_DRIVER], input=fmt(words), capture_output=True, text=True,
                             check=True).stdout.strip().split("\n")
        out["adp%d_meta" % k] = np.array([dim, degree] + reps)
        out["adp%d_vec" % k], out["adp%d_buf" % k] = vec, buf
        out["adp%d_write" % k] = np.array(res[0].split(), dtype=float)
        out["adp%d_after_read" % k] = np.array(res[1].split(), dtype=float)
        out["adp%d_checkpoint" % k] = np.array(
            " ".join(l for l in res[2:] if not l.startswith("EVENT")).split(), dtype=float)
        out["adp%d_events" % k] = np.array([l[6:] for l in res if l.startswith("EVENT")])
    return out


def run_grid_cases():
    """make_grid of both solvers (nonlinear_elasticity.cc:169-285, linear_elasticity.cc:79-187):
    per (solver, dim, scenario): [volume, dim, refinements, interface id, clamped id, z-clamp id],
    repetitions, box corners, and the boundary id each colorized face (x-,x+,y-,y+,z-,z+) gets."""
    out = {}
    k = 0
    for solver in ("nl", "lin"):
        for dim in (2, 3):
            for scenario, flap in (("FSI3", 0.0), ("PF", 0.0), ("PF", 0.25)):
                txt = subprocess.run([GRID_DRIVER, solver, str(dim), scenario, repr(flap)],
                                     capture_output=True, text=True, check=True).stdout.split("\n")
                out["grid%02d_key" % k] = np.array([solver, str(dim), scenario, repr(flap)])
                out["grid%02d_head" % k] = np.array([float(txt[0])] + [float(x) for x in txt[1].split()])
                out["grid%02d_reps" % k] = np.array(txt[2].split(), dtype=np.int64)
                out["grid%02d_p0" % k] = np.array(txt[3].split(), dtype=float)
                out["grid%02d_p1" % k] = np.array(txt[4].split(), dtype=float)
                out["grid%02d_face_ids" % k] = np.array(txt[5].split(), dtype=np.int64)
                k += 1
    out["n_grid"] = np.array(k)
    return out


def prm_cases():
    base = open(os.path.join(HERE, "parameters_nonlinear_fsi3.prm")).read()
    linear = """# linear model, conservative data, body force, other scenario
subsection Time
  set End time        = 2.5
  set Time step size  = 0.005   # trailing comment
  set Output interval = 25
  set Output folder   = out
end
subsection Discretization
  set Polynomial degree = 1
  set theta             = 0.6
end
subsection System properties
  set Shear modulus   = 1.2e6
  set Poisson's ratio = 0.25
  set rho             = 3000
  set body forces     = 0.0, -9.81, 0.5
end
subsection Solver
  set Model                    = linear
  set Solver type              = CG
  set Max iteration multiplier = 1.5
end
subsection precice configuration
  set Scenario         = PF
  set Read data name   = Force-Data
  set Flap location    = 1.5
  set Participant name = Flap
end
"""
    cases = [base, "", linear,
             base.replace("set rho ", "set density "),                        # undeclared entry
             base.replace("subsection Solver", "subsection Linear solver"),   # undeclared subsection
             base.replace("= 0.4", "= 0.7"),                                  # nu outside [-1, 0.5]
             base.replace("= neo-Hookean", "= hyperelastic"),                 # not in the selection
             base.replace("= Stress", "= Pressure"),                          # neither Stress nor Force
             base.replace("set beta                = 0.25", "set beta                = 0.7"),
             base.replace("Output interval       = 10", "Output interval       = -1"),
             linear.replace("0.0, -9.81, 0.5", "0.0, -9.81")]                # two body-force entries
    return cases


def run_prm_cases():
    import tempfile
    out = {}
    cases = prm_cases()
    for k, text in enumerate(cases):
        with tempfile.NamedTemporaryFile("w", suffix=".prm", delete=False) as f:
            f.write(text)
        r = subprocess.run([PRM_DRIVER, f.name], capture_output=True, text=True)
        os.unlink(f.name)
        out["prm%02d_text" % k] = np.array(text)
        out["prm%02d_exit" % k] = np.array(r.returncode)
        out["prm%02d_out" % k] = np.array(r.stdout if r.returncode == 0 else "")
    out["n_prm"] = np.array(len(cases))
    return out


def run_loop_cases():
    """Event order of the reference's own run() loops (nonlinear_elasticity.cc:96-167,
    linear_elasticity.cc:632-716) under a scripted coupling scheme:
    (solver, windows, sub-iterations, output interval, solver dt, preCICE window size)."""
    cases = [("nl", 2, 1, 1, 0.01, 0.01), ("nl", 2, 3, 1, 0.01, 0.01), ("nl", 3, 2, 2, 0.01, 0.01),
             ("lin", 3, 1, 1, 0.01, 0.01), ("lin", 2, 2, 1, 0.01, 0.01),
             ("nl", 1, 1, 1, 0.01, 0.02), ("lin", 1, 1, 1, 0.01, 0.02)]
    out = {}
    for k, c in enumerate(cases):
        r = subprocess.run([RUN_DRIVER] + [str(x) for x in c], capture_output=True, text=True)
        out["loop%d_case" % k] = np.array([str(x) for x in c])
        out["loop%d_exit" % k] = np.array(r.returncode)
        out["loop%d_events" % k] = np.array(r.stdout.strip().split("\n"))
    out["n_loop"] = np.array(len(cases))
    return out


def run_constraint_cases():
    """Dirichlet sets the reference requests (make_constraints nonlinear_elasticity.cc:1094-1150 for
    Newton iterations 0, 1, 2; linear_elasticity.cc:429-446): lines `call boundary_id mask-bits`."""
    out = {}
    cases = [("nl", 2, 0), ("nl", 2, 1), ("nl", 2, 2), ("nl", 3, 0), ("nl", 3, 1), ("nl", 3, 2),
             ("lin", 2), ("lin", 3)]
    for k, c in enumerate(cases):
        r = subprocess.run([CONSTRAINTS_DRIVER] + [str(x) for x in c], capture_output=True, text=True,
                           check=True)
        out["cst%d_case" % k] = np.array([str(x) for x in c])
        out["cst%d_calls" % k] = np.array(r.stdout.strip().split("\n") if r.stdout.strip() else [])
    out["n_cst"] = np.array(len(cases))
    return out


def run_solver_setup_cases():
    """What the reference hands to SolverControl / the preconditioner (nonlinear_elasticity.cc:
    1153-1211, linear_elasticity.cc:525-575): (solver, type, n_dofs, multiplier, tol_lin, |rhs|)."""
    cases = [("nl", "CG", 2081667, 1.0, 1e-6, 350.5), ("nl", "CG", 7, 1.5, 1e-6, 2.0),
             ("nl", "CG", 518, 2.0, 1e-8, 0.0), ("nl", "Direct", 100, 1.0, 1e-6, 3.0),
             ("lin", "CG", 7, 1.5, 0.0, 0.0), ("lin", "CG", 51171075, 1.0, 0.0, 0.0),
             ("lin", "Direct", 7, 1.5, 0.0, 0.0)]
    out = {}
    for k, c in enumerate(cases):
        r = subprocess.run([SOLVER_DRIVER] + [repr(x) if isinstance(x, float) else str(x) for x in c],
                           capture_output=True, text=True, check=True)
        out["slv%d_case" % k] = np.array([str(x) for x in c])
        out["slv%d_lines" % k] = np.array(r.stdout.strip().split("\n"))
    out["n_slv"] = np.array(len(cases))
    return out


def generate():
    out = {}
    out.update(run_solver_setup_cases())
    out.update(run_constraint_cases())
    out.update(run_loop_cases())
    out.update(run_prm_cases())
    out.update(run_grid_cases())
    out.update(run_adapter_cases())
    out.update(run_update_cases())
    out.update(run_newton_scripts())
    lcases = linear_cases()
    for k, c in enumerate(lcases):
        out.update(run_linear_case(k, *c))
    out["n_linear"] = np.array(len(lcases))
    # ---- cell assembly: tangent, residual, Neumann term of one cell -------------------------
    cases = assembly_cases()
    for k, c in enumerate(cases):
        out.update(run_assembly_case(k, *c))
    dcases = distorted_assembly_cases()
    for k, c in enumerate(dcases):
        out.update(run_distorted_assembly_case(len(cases) + k, *c))
    out["n_assembly"] = np.array(len(cases) + len(dcases))
    # ---- material -------------------------------------------------------------------------
    rows = []
    for k, (dim, mu, nu, J, bv) in enumerate(material_cases()):
        n = dim * (dim + 1) // 2
        txt = run(["material", dim, mu, nu, 1000.0, J] + bv).split("\n")
        psi = float(txt[0])
        tau = np.array(txt[1].split(), dtype=float)
        Jc = np.array([l.split() for l in txt[2:2 + n]], dtype=float)
        asym = float(txt[2 + n])
        assert tau.shape == (n,) and Jc.shape == (n, n)
        out["mat%02d_in" % k] = np.array([dim, mu, nu, J] + bv)
        out["mat%02d_psi" % k] = np.array(psi)
        out["mat%02d_tau" % k] = tau
        out["mat%02d_Jc" % k] = Jc
        out["mat%02d_asym" % k] = np.array(asym)
        rows.append(k)
    out["n_material"] = np.array(len(rows))
    # ---- Postprocessor: u and grad_x u = A (I + A)^-1 of the linear field u = A X ------------
    rng = np.random.RandomState(7)
    for dim in (2, 3):
        A = 0.08 * rng.uniform(-1, 1, (dim, dim))
        g = A @ np.linalg.inv(np.eye(dim) + A)
        X = rng.uniform(0, 1, (5, dim))
        lines = []
        for x in X:
            u = A @ x
            lines.append(" ".join(repr(float(v)) for v in list(u) + list(g.reshape(-1))))
        txt = run(["strain", dim], stdin="\n".join(lines) + "\n").split("\n")
        out["pp%d_A" % dim] = A
        out["pp%d_X" % dim] = X
        out["pp%d_names" % dim] = np.array(txt[0].split())
        out["pp%d_out" % dim] = np.array([l.split() for l in txt[1:1 + len(X)]], dtype=float)
    # ---- Time -----------------------------------------------------------------------------------
    tcases = [(1.0, 0.01, 7, 0.03), (10.0, 0.005, 13, 0.045), (0.5, 0.1, 3, 0.1), (2.0, 0.0025, 400, 0.9975)]
    tout = []
    for t_end, dt, n_inc, t_reset in tcases:
        txt = run(["time", t_end, dt, n_inc, t_reset]).split("\n")
        a, b = txt[0].split(), txt[1].split()
        tout.append([t_end, dt, n_inc, t_reset, float(a[0]), float(a[1]), float(b[0]), float(b[1])])
    out["time_cases"] = np.array(tout)
    return out


if __name__ == "__main__":
    build()
    data = generate()
    if "--check" in sys.argv:
        old = np.load(OUT)
        bad = [k for k in data if not np.array_equal(np.asarray(data[k]), old[k])]
        print("differences:", bad)
        sys.exit(1 if bad else 0)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, "with", len(data), "arrays")
