"""Sampled-row parity at BASELINE.json's full sizes (cfg3: 2.08 M DoFs, cfg4: 51 M DoFs), where the
CPU oracle cannot assemble the whole matrix inside a test.

A global matrix row of DoF i (and the residual entry i) only receives contributions from the cells
that contain i's node (`distribute_local_to_global`, nonlinear_elasticity.cc:760-774;
`stiffness_matrix.add`, linear_elasticity.cc:327-334). So the oracle is run on the SUB-MESH made of
all cells touching the sampled nodes - same cell order, same constraint flags, same interface
faces, same state restricted to its DoFs - and its rows at the sampled DoFs are complete: they are
what the oracle would produce on the full mesh. Rows of the other sub-mesh DoFs are incomplete and
never compared.
"""
from types import SimpleNamespace

import numpy as np


def node_dofs(problem):
    """[n_nodes, dim] DoF ids per node (x-, y-, z-component), ordered by the x-component id.
    FESystem(FE_Q(p), dim) with p <= 2 has one DoF per entity and component: local i = a*dim + c."""
    dim = problem.dim
    nd = np.asarray(problem.mesh.cell_dofs).reshape(-1, dim)
    _, first = np.unique(nd[:, 0], return_index=True)
    return nd[first]


def sample_nodes(problem, n_per_class=24, seed=7, cut_parts=(2, 4, 8), axis=1):
    """Deterministic node sample covering the row classes that differ in the assembly: clamped
    face, z-clamped faces, interface faces, edges / corners of the box, the node planes where
    slab partitions into `cut_parts` ranks cut (plus the plane after), interior."""
    nd = node_dofs(problem)
    x = np.asarray(problem.mesh.support_points)[nd[:, 0]]
    p0, p1 = np.array(problem.mesh.p0), np.array(problem.mesh.p1)
    tol = 1e-9 * (p1 - p0).max()
    on_lo = np.abs(x - p0) < tol
    on_hi = np.abs(x - p1) < tol
    n_faces = (on_lo | on_hi).sum(axis=1)
    classes = [n_faces == 0, n_faces == 1, n_faces == 2, n_faces >= 3]
    for d in range(problem.dim):
        classes += [on_lo[:, d], on_hi[:, d]]
    reps = problem.mesh.reps
    p = problem.degree
    h_ax = (p1[axis] - p0[axis]) / (reps[axis] * p)            # node-plane spacing along the axis
    plane = np.rint((x[:, axis] - p0[axis]) / h_ax).astype(np.int64)
    for parts in cut_parts:
        if reps[axis] % parts == 0:
            for r in range(1, parts):
                cut = r * (reps[axis] // parts) * p
                classes += [plane == cut, plane == cut + 1, plane == cut - 1]
    rng = np.random.RandomState(seed)
    chosen = set()
    for mask in classes:
        idx = np.flatnonzero(mask)
        if len(idx):
            chosen.update(rng.choice(idx, size=min(n_per_class, len(idx)), replace=False).tolist())
    return nd[np.array(sorted(chosen))]                         # [n_sample, dim]


def sub_problem(problem, sample):
    """Sub-mesh of all cells touching the sampled nodes. Returns (sub, dof_map, cells): `sub` has the
    attributes the oracle binding reads; dof_map[sub dof] = global dof (ascending, so column order
    is preserved); cells = global indices of the sub-mesh cells (ascending = same scatter order)."""
    dim = problem.dim
    mesh = problem.mesh
    dpc = mesh.dofs_per_cell
    nv = 1 << dim
    cd = np.asarray(mesh.cell_dofs).reshape(-1, dpc)
    touch = np.isin(cd[:, 0::dim], sample[:, 0]).any(axis=1)
    cells = np.flatnonzero(touch)
    dof_map, inv = np.unique(cd[cells].reshape(-1), return_inverse=True)
    sub_mesh = SimpleNamespace(
        n_cells=len(cells), n_dofs=len(dof_map), dofs_per_cell=dpc,
        cell_dofs=inv.astype(np.int32),
        cell_vertices=np.asarray(mesh.cell_vertices).reshape(-1, nv * dim)[cells].reshape(-1).copy())
    keep_f = np.isin(problem.iface_cell, cells)
    keep_n = np.isin(problem.iface_dofs[0], dof_map)
    sub = SimpleNamespace(
        dim=dim, degree=problem.degree, model=problem.model, params=problem.params, mesh=sub_mesh,
        n_dofs=len(dof_map), constrained=np.asarray(problem.constrained)[dof_map].copy(),
        iface_cell=np.searchsorted(cells, problem.iface_cell[keep_f]).astype(np.int32),
        iface_face_no=problem.iface_face_no[keep_f].astype(np.int32),
        iface_dofs=np.searchsorted(dof_map, problem.iface_dofs[:, keep_n]).astype(np.int32),
        n_iface_nodes=int(keep_n.sum()))
    return sub, dof_map, cells


def rows_of(csr, sub_rows, dof_map):
    """(rowptr, global cols, vals) of the given sub-mesh rows of a scipy CSR in sub numbering."""
    A = csr[sub_rows]
    A.sort_indices()
    return A.indptr.astype(np.int64), dof_map[A.indices].astype(np.int32), A.data.copy()


def traction_vector(problem, buf):
    """What format_precice_to_deal (adapter.h:421-443) scatters into the DoF vector."""
    v = np.zeros(problem.n_dofs)
    b = np.asarray(buf).reshape(-1, problem.dim)
    for c in range(problem.dim):
        v[problem.iface_dofs[c]] = b[:, c]
    return v


def oracle_nl_rows(orc, problem, sample, u, du, v_old, a_old, traction_buf):
    """Oracle tangent rows + residual entries at the sampled DoFs for the given state (the state
    update_acceleration + assemble_system see, nonlinear_elasticity.cc:444-446)."""
    sub, dof_map, cells = sub_problem(problem, sample)
    o = orc.Oracle(sub)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u[dof_map])
    o.set(orc.NL_SOLUTION_DELTA, du[dof_map])
    o.set(orc.NL_VELOCITY_OLD, v_old[dof_map])
    o.set(orc.NL_ACCELERATION_OLD, a_old[dof_map])
    o.set(orc.NL_EXTERNAL_STRESS, traction_vector(problem, traction_buf)[dof_map])
    o.nl_update_acceleration()
    o.nl_assemble_system()
    rows = sample.reshape(-1)
    sub_rows = np.searchsorted(dof_map, rows)
    rp, col, val = rows_of(o.csr(orc.MAT_TANGENT), sub_rows, dof_map)
    return {"rows": rows.astype(np.int32), "rowptr": rp, "col": col, "val": val,
            "rhs": o.get(orc.NL_SYSTEM_RHS)[sub_rows], "n_sub_cells": len(cells)}


def oracle_lin_rows(orc, problem, sample):
    """Oracle rows of K, M and the constrained system matrix (linear_elasticity.cc:248-374,
    426-451) at the sampled DoFs."""
    sub, dof_map, cells = sub_problem(problem, sample)
    o = orc.Oracle(sub)
    o.lin_assemble_system()
    o.lin_assemble_rhs()          # builds system_matrix (copy + apply_boundary_values)
    rows = sample.reshape(-1)
    sub_rows = np.searchsorted(dof_map, rows)
    out = {"rows": rows.astype(np.int32), "n_sub_cells": len(cells)}
    for name, which in (("K", orc.MAT_STIFFNESS), ("M", orc.MAT_MASS), ("A", orc.MAT_SYSTEM)):
        rp, col, val = rows_of(o.csr(which), sub_rows, dof_map)
        out["rowptr"], out["col"], out[name] = rp, col, val
    return out


def assert_rows_close(got, ref_rowptr, ref_col, ref_val, tol=1e-12):
    rp, col, val = got
    assert np.array_equal(rp, ref_rowptr), "row lengths differ"
    assert np.array_equal(col, ref_col), "column indices differ"
    n = len(rp) - 1
    rowmax = np.maximum.reduceat(np.abs(ref_val), ref_rowptr[:-1])
    err = np.abs(val - ref_val) / np.repeat(np.maximum(rowmax, 1e-300), np.diff(ref_rowptr))
    assert err.max() <= tol, "max row-relative error %.3e over %d rows" % (err.max(), n)
    return float(err.max())
