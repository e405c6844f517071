"""The generic-degree kernels (dealii_adapter_b200/csrc/assemble_nl_generic.cuh: the reference's
shipped parameter files ask for polynomial degree 3 and 4, parameters.prm:21,
nonlinear_elasticity.prm:24) were written in a session without GPU access. The SAME source is
therefore also compiled with g++ against the CUDA stand-in of tests/cuda_emu/ (a CTA = CPU threads +
a barrier) and checked here against the oracle: element tangents (lower node blocks) and element
residuals of assemble_system_tangent_residual_one_cell (nonlinear_elasticity.cc:872-1036), and the
assembled right-hand side including assemble_neumann_contribution_one_cell (:791-859).
The GPU run of the same kernels is tests/test_gpu_zzz_high_degree.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import nl_params, smooth_field
from dealii_adapter_b200.problem import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "cuda_emu")


@pytest.fixture(scope="module")
def emu(native_libs):
    out = os.path.join(EMU, "_build", "libgf_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    csrc = os.path.join(ROOT, "dealii_adapter_b200", "csrc")
    srcs = [os.path.join(EMU, "emu_kernels.cpp"), os.path.join(EMU, "cuda_runtime.h")] + \
        [os.path.join(csrc, f) for f in ("assemble_nl_generic.cuh", "assemble_general.cuh",
                                         "direct_band.cuh", "rcm.h", "constraints.cuh",
                                         "emu_compat.cuh",
                                         "kernel_utils.cuh", "nl_material.cuh", "fe_tables_host.h",
                                         "fe_basis.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-Wall",
                        "-Wno-unknown-pragmas", "-DGF_CUDA_EMULATION", "-I", EMU, "-I", csrc,
                        "-I", os.path.join(ROOT, "include"), "-o", out, srcs[0], "-lpthread"],
                       check=True)
    lib = C.CDLL(out)
    lib.emu_nl_cells.restype = C.c_int
    lib.emu_nl_cells.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 6 + \
        [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    lib.emu_nl_faces.restype = C.c_int
    lib.emu_nl_faces.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_uint, C.c_void_p]
    lib.emu_linear_general.restype = C.c_int
    lib.emu_linear_general.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                       C.c_double, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + \
        [C.c_uint] + [C.c_void_p] * 4
    lib.emu_postprocess_general.restype = C.c_int
    lib.emu_postprocess_general.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_uint, C.c_void_p]
    lib.emu_loc_of.restype = C.c_int
    lib.emu_loc_of.argtypes = [C.c_int, C.c_int, C.c_void_p]
    return lib


def device_view(prob, emu):
    """What gf_create derives from the mesh description: internal node numbering (any bijection
    will do here), cell -> node table in hierarchical local order, affine geometry records."""
    dim, p, mesh = prob.dim, prob.degree, prob.mesh
    npc = (p + 1) ** dim
    loc_of = np.zeros(npc * dim, dtype=np.int32)
    assert emu.emu_loc_of(dim, p, loc_of.ctypes.data) == npc
    cd = mesh.cell_dofs.reshape(mesh.n_cells, npc * dim)
    xdofs = cd[:, loc_of[0::dim]]                         # [cell, a]: global x-dof of local node a
    ids, cell_nodes = np.unique(xdofs.reshape(-1), return_inverse=True)
    cell_nodes = cell_nodes.reshape(mesh.n_cells, npc).astype(np.int32)
    # internal dof (node, c) -> caller's global dof
    i2e = np.zeros((len(ids), dim), dtype=np.int64)
    for c in range(dim):
        i2e[cell_nodes.reshape(-1), c] = cd[:, loc_of[c::dim]].reshape(-1)
    cv = mesh.cell_vertices.reshape(mesh.n_cells, 1 << dim, dim)
    h = cv[:, -1, :] - cv[:, 0, :]
    geom = np.zeros((mesh.n_cells, dim * dim + 1))
    for d in range(dim):
        geom[:, d * dim + d] = 1.0 / h[:, d]
    geom[:, -1] = np.prod(h, axis=1)
    return loc_of, cell_nodes, i2e.reshape(-1), geom


CASES = [(2, 3, [3, 2]), (2, 4, [2, 2]), (2, 5, [2, 1]), (3, 3, [2, 1, 2]), (2, 2, [3, 2]),
         (3, 2, [2, 2, 1])]


def is_affine(prob):
    dim = prob.dim
    v = prob.mesh.cell_vertices.reshape(prob.mesh.n_cells, 1 << dim, dim)
    last = v[:, 0] + sum(v[:, 1 << d] - v[:, 0] for d in range(dim))
    h = np.abs(v[:, -1] - v[:, 0]).max()
    return np.abs(last - v[:, -1]).max() < 1e-6 * h


def iface_lists(prob):
    """cells with interface faces, their faces in ascending face number (as gf_create groups them)"""
    icells = np.unique(prob.iface_cell)
    order = np.lexsort((prob.iface_face_no, prob.iface_cell))
    face_no = prob.iface_face_no[order].astype(np.int32)
    face_ptr = np.concatenate([np.searchsorted(prob.iface_cell[order], icells),
                               [len(order)]]).astype(np.int32)
    return icells.astype(np.int32), face_ptr, face_no


@pytest.mark.parametrize("dim,p,reps", CASES)
@pytest.mark.parametrize("block,general", [(32, False), (96, False), (64, True)])
def test_generic_cell_and_face_kernels_match_oracle(emu, dim, p, reps, block, general):
    """general = True: distorted (non-affine) cells, Jacobians per quadrature point (AFFINE = false)"""
    from oracle import oracle_py as orc
    from helpers import distort_mesh
    prm = nl_params(poly_degree=p, body_force=(1.5, -9.81, 0.5), scenario="PF")
    prob = make_problem(prm, dim, reps=reps, numbering="cellwise")
    prob.constrained[:] = 0
    if general:
        distort_mesh(prob, 0.12, seed=dim * 10 + p)
        assert not is_affine(prob)
    mesh = prob.mesh
    npc, dpc, nc = (p + 1) ** dim, dim * (p + 1) ** dim, mesh.n_cells
    loc_of, cell_nodes, i2e, geom = device_view(prob, emu)
    L = (np.array(mesh.p1) - np.array(mesh.p0)).min()
    u = smooth_field(prob, 0.05 * L, seed=5)
    acc = smooth_field(prob, 40.0, seed=6)
    rng = np.random.RandomState(7)
    stress = 2000.0 * rng.uniform(-1, 1, prob.n_dofs)
    o = orc.Oracle(prob, n_threads=1)
    kappa = (2.0 * prm.mu * (1.0 + prm.nu)) / (3.0 * (1.0 - 2.0 * prm.nu))
    alpha_1 = 1.0 / (prm.beta * prm.delta_t ** 2)
    params = np.array([kappa, prm.mu, prm.rho, alpha_1] + list(prm.body_force), dtype=np.float64)
    u_i, acc_i, stress_i = (np.ascontiguousarray(v[i2e]) for v in (u, acc, stress))
    ke = np.full((nc, dpc, dpc), np.nan)
    re = np.full((nc, dpc), np.nan)
    grid = 2 if nc > 2 else 1          # grid-stride loop over the cells
    verts = np.ascontiguousarray(mesh.cell_vertices, dtype=np.float64)
    geom_p, verts_p = (None, verts.ctypes.data) if general else (geom.ctypes.data, None)
    err = emu.emu_nl_cells(dim, p, nc, cell_nodes.ctypes.data, geom_p, verts_p, u_i.ctypes.data,
                           acc_i.ctypes.data, params.ctypes.data, grid, block, ke.ctypes.data,
                           re.ctypes.data)
    assert err == 0
    cd = mesh.cell_dofs.reshape(nc, dpc)
    node = np.arange(dpc) // dim
    lower = node[:, None] >= node[None, :]              # node blocks b <= a
    for cell in range(nc):
        K, r = o.nl_cell(cell, u[cd[cell]], acc[cd[cell]])
        K_int = K[np.ix_(loc_of, loc_of)]               # internal order (a, c) -> a * dim + c
        scale = np.abs(K).max()
        assert np.abs(ke[cell][lower] - K_int[lower]).max() <= 1e-12 * scale
        assert np.isnan(ke[cell][~lower]).all()         # the upper blocks are never written
        assert np.abs(re[cell] - r[loc_of]).max() <= 1e-12 * np.abs(r).max()
    # interface faces on top (K2g), then the assembled right-hand side against the oracle's
    cell_list, face_ptr, face_no = iface_lists(prob)
    err = emu.emu_nl_faces(dim, p, len(cell_list), cell_list.ctypes.data, face_ptr.ctypes.data,
                           face_no.ctypes.data, cell_nodes.ctypes.data, geom_p, verts_p,
                           u_i.ctypes.data, stress_i.ctypes.data, block, re.ctypes.data)
    assert err == 0
    rhs = np.zeros(prob.n_dofs)
    np.add.at(rhs, cd[:, loc_of].reshape(-1), re.reshape(-1))
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    o.set(orc.NL_ACCELERATION, acc)
    o.set(orc.NL_EXTERNAL_STRESS, stress)
    o.nl_assemble_system()
    ref = o.get(orc.NL_SYSTEM_RHS)
    assert np.abs(rhs - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("dim,p,reps", [(2, 1, [3, 3]), (2, 2, [3, 2]), (2, 3, [2, 2]), (3, 1, [2, 2, 2]),
                                        (3, 2, [2, 1, 2]), (3, 3, [1, 2, 1])])
def test_general_cell_linear_and_output_kernels_match_oracle(emu, dim, p, reps):
    """assemble_general.cuh on distorted meshes: stiffness, mass, body-force load, consistent
    interface load (linear_elasticity.cc:289-323, 340-345, 358-373, 483-520) and the output fields
    (DataOut patches + Postprocessor) against the oracle."""
    import scipy.sparse as sp
    from oracle import oracle_py as orc
    from helpers import distort_mesh, lin_params
    prm = lin_params(poly_degree=p, body_force=(0.4, -9.81, 0.3), scenario="PF")
    prob = make_problem(prm, dim, reps=reps, numbering="cellwise")
    prob.constrained[:] = 0
    distort_mesh(prob, 0.12, seed=dim * 10 + p)
    assert not is_affine(prob)
    mesh = prob.mesh
    npc, dpc, nc, n = (p + 1) ** dim, dim * (p + 1) ** dim, mesh.n_cells, prob.n_dofs
    loc_of, cell_nodes, i2e, _ = device_view(prob, emu)
    verts = np.ascontiguousarray(mesh.cell_vertices, dtype=np.float64)
    rng = np.random.RandomState(11)
    stress = 2000.0 * rng.uniform(-1, 1, n)
    stress_i = np.ascontiguousarray(stress[i2e])
    cell_list, face_ptr, face_no = iface_lists(prob)
    ke, me = np.full((nc, dpc, dpc), np.nan), np.full((nc, npc, npc), np.nan)
    re_body, re_face = np.full((nc, dpc), np.nan), np.zeros((nc, dpc))
    bf = np.array(prm.body_force, dtype=np.float64)
    err = emu.emu_linear_general(dim, p, nc, verts.ctypes.data, prm.lam, prm.mu, prm.rho,
                                 bf.ctypes.data, len(cell_list), cell_list.ctypes.data,
                                 face_ptr.ctypes.data, face_no.ctypes.data, cell_nodes.ctypes.data,
                                 stress_i.ctypes.data, 64, ke.ctypes.data, me.ctypes.data,
                                 re_body.ctypes.data, re_face.ctypes.data)
    assert err == 0
    cd = mesh.cell_dofs.reshape(nc, dpc)[:, loc_of]          # [cell, a * dim + c] -> global dof
    K, M = np.zeros((n, n)), np.zeros((n, n))
    for c in range(nc):
        K[np.ix_(cd[c], cd[c])] += ke[c]
        for cc in range(dim):                                # M = m_ab delta_cd
            M[np.ix_(cd[c, cc::dim], cd[c, cc::dim])] += me[c]
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    K_o, M_o = o.csr(orc.MAT_STIFFNESS).toarray(), o.csr(orc.MAT_MASS).toarray()
    assert np.abs(K - K_o).max() <= 1e-12 * np.abs(K_o).max()
    assert np.abs(M - M_o).max() <= 1e-12 * np.abs(M_o).max()
    body = np.zeros(n)
    np.add.at(body, cd.reshape(-1), re_body.reshape(-1))
    ref = o.get(orc.LIN_BODY_FORCE)
    assert np.abs(body - ref).max() <= 1e-12 * np.abs(ref).max()
    load = np.zeros(n)
    np.add.at(load, cd.reshape(-1), re_face.reshape(-1))
    o.set(orc.LIN_STRESS, stress)
    o.lin_assemble_rhs()              # old_stress <- consistent loading + body force (:405-409)
    ref = o.get(orc.LIN_OLD_STRESS)
    assert np.abs(load + body - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(load).max() > 0.1 * np.abs(ref).max()      # the interface term is in there
    # output fields of a smooth displacement on the distorted mesh
    L = (np.array(mesh.p1) - np.array(mesh.p0)).min()
    u = smooth_field(prob, 0.05 * L, seed=3)
    o.set(orc.LIN_DISPLACEMENT, u)
    _, fld_o = o.postprocess(orc.LIN_DISPLACEMENT)
    u_i = np.ascontiguousarray(u[i2e])
    fld = np.full((nc, npc, dim + dim * dim), np.nan)
    err = emu.emu_postprocess_general(dim, p, nc, cell_nodes.ctypes.data, verts.ctypes.data,
                                      u_i.ctypes.data, 64, fld.ctypes.data)
    assert err == 0
    assert np.abs(fld - fld_o).max() <= 1e-12 * max(1.0, np.abs(fld_o).max())


def test_constraint_line_kernels(emu):
    """constraints.cuh: x <- C x on the hanging dofs, y <- C^T y (+ zero) against scipy, on the
    hanging-node mesh of tests/helpers.py (Q2: two hanging nodes, three masters each)."""
    from helpers import constraint_matrix, hanging_node_problem, lin_params
    emu.emu_lines.restype = None
    emu.emu_lines.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 6
    for p in (1, 2):
        prob = hanging_node_problem(lin_params(poly_degree=p), p)
        dof, ptr, master, weight = prob.extra["constraint_lines"]
        Cm = constraint_matrix(prob)
        # transposed lists as constraints.cu builds them: masters ascending, lines in order
        by_master = {}
        for k, s_ in enumerate(dof):
            for j in range(ptr[k], ptr[k + 1]):
                by_master.setdefault(int(master[j]), []).append((int(s_), float(weight[j])))
        mdof = np.array(sorted(by_master), dtype=np.int32)
        mptr = np.concatenate([[0], np.cumsum([len(by_master[m]) for m in mdof])]).astype(np.int64)
        slave = np.array([s_ for m in mdof for s_, _ in by_master[m]], dtype=np.int32)
        mweight = np.array([w for m in mdof for _, w in by_master[m]])
        rng = np.random.RandomState(p)
        x, y = rng.uniform(-1, 1, prob.n_dofs), rng.uniform(-1, 1, prob.n_dofs)
        x_ref = x.copy()
        x_ref[dof] = 0.0
        x_ref = Cm @ x_ref                     # masters keep their values, hanging dofs interpolate
        y_ref = Cm.T @ y                       # hanging entries of C^T y are empty rows -> 0
        emu.emu_lines(len(dof), dof.ctypes.data, ptr.ctypes.data, master.ctypes.data,
                      weight.ctypes.data, len(mdof), mdof.ctypes.data, mptr.ctypes.data,
                      slave.ctypes.data, mweight.ctypes.data, x.ctypes.data, y.ctypes.data)
        assert np.abs(x - x_ref).max() < 1e-15
        assert np.abs(y - y_ref).max() < 1e-15 and np.all(y[dof] == 0.0)


# ------------------------------------------------------------------------------------------------
# 'Solver type = Direct' (parameters.prm:43): csrc/direct_band.cuh + rcm.h, same source on the CPU
# ------------------------------------------------------------------------------------------------
def block_row_format(A, dim):
    """scipy CSR (dof = node * dim + c) -> the library's block-row arrays (gf_context.h)."""
    n_nodes = A.shape[0] // dim
    brow_ptr, bcol, val_ptr, vals = [0], [], [0], []
    Ad = A.tocsr()
    for a in range(n_nodes):
        cols = np.unique(np.concatenate([Ad.indices[Ad.indptr[a * dim + r]:Ad.indptr[a * dim + r + 1]] // dim
                                         for r in range(dim)]))
        stride = (len(cols) * dim + 1) & ~1
        rows = np.zeros((dim, stride))
        dense = Ad[a * dim:(a + 1) * dim, :].toarray()
        for k, bnode in enumerate(cols):
            rows[:, k * dim:(k + 1) * dim] = dense[:, bnode * dim:(bnode + 1) * dim]
        bcol += list(cols)
        brow_ptr.append(len(bcol))
        vals.append(rows.reshape(-1))
        val_ptr.append(val_ptr[-1] + dim * stride)
    return (np.array(brow_ptr, dtype=np.int32), np.array(val_ptr, dtype=np.int64),
            np.array(bcol, dtype=np.int32), np.concatenate(vals))


@pytest.mark.parametrize("dim,p,reps,numbering", [(2, 2, [9, 3], "lexicographic"),
                                                  (2, 3, [6, 2], "cellwise"),
                                                  (3, 1, [3, 7, 2], "lexicographic"),
                                                  (3, 2, [2, 4, 1], "cellwise")])
def test_band_cholesky_solves_the_tangent_system(emu, dim, p, reps, numbering):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import oracle_py as orc
    emu.emu_direct_solve.restype = C.c_int
    emu.emu_direct_solve.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 7
    prob = make_problem(nl_params(poly_degree=p), dim, reps=reps, numbering=numbering)
    o = orc.Oracle(prob)
    L = (np.array(prob.mesh.p1) - np.array(prob.mesh.p0)).min()
    o.set(orc.NL_TOTAL_DISPLACEMENT, smooth_field(prob, 0.03 * L, seed=2))
    o.format_precice_to_deal(np.tile([1500.0, -300.0, 200.0][:dim], prob.n_iface_nodes),
                             orc.NL_EXTERNAL_STRESS)
    o.nl_assemble_system()
    rowptr, col = o.pattern()
    n = prob.n_dofs
    A = sp.csr_matrix((o.values(orc.MAT_TANGENT), col, rowptr), shape=(n, n))
    b = o.get(orc.NL_SYSTEM_RHS)
    brow_ptr, val_ptr, bcol, val = block_row_format(A, dim)
    x = np.zeros(n)
    w = np.zeros(2, dtype=np.int64)
    info = emu.emu_direct_solve(n // dim, dim, brow_ptr.ctypes.data, val_ptr.ctypes.data,
                                bcol.ctypes.data, val.ctypes.data, b.ctypes.data, x.ctypes.data,
                                w.ctypes.data)
    assert info == 0
    ref = spla.spsolve(A.tocsc(), b)
    assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(A @ x - b).max() <= 1e-10 * np.abs(b).max()
    assert w[0] <= w[1]                 # the better of RCM and the given numbering
    # an indefinite matrix is reported, not factorised silently
    val_bad = val.copy()
    val_bad[0] = -abs(val_bad[0])       # first diagonal entry (node 0 couples to itself first)
    assert bcol[0] == 0
    info = emu.emu_direct_solve(n // dim, dim, brow_ptr.ctypes.data, val_ptr.ctypes.data,
                                bcol.ctypes.data, val_bad.ctypes.data, b.ctypes.data, x.ctypes.data,
                                w.ctypes.data)
    assert info > 0
