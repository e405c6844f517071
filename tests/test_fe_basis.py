"""FE_Q(p) for every degree the parameter files ask for (parameters.prm:21 = 3,
nonlinear_elasticity.prm:24 = 4): csrc/fe_basis.h - shared by the device library's table builders
and the host mesh - against the independent numpy transcription in tests/ref_formulas.py."""
import ctypes as C

import numpy as np
import pytest

import ref_formulas as rf
from dealii_adapter_b200.mesh import StructuredMesh, host_lib


def _lib():
    lib = host_lib()
    lib.gfh_fe_support_points.argtypes = [C.c_int, C.c_void_p]
    lib.gfh_fe_basis_eval.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gfh_fe_numbering.restype = C.c_int
    lib.gfh_fe_numbering.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 8])
def test_support_points_and_basis(p):
    lib = _lib()
    pts = np.zeros(p + 1)
    lib.gfh_fe_support_points(p, pts.ctypes.data)
    assert np.abs(pts - rf.support_points_1d(p)).max() < 1e-15
    assert pts[0] == 0.0 and pts[-1] == 1.0 and np.all(np.diff(pts) > 0)
    assert np.abs(pts + pts[::-1] - 1.0).max() < 1e-16          # symmetric
    x = np.concatenate([np.linspace(0, 1, 23), rf.gauss01(p + 2)[0], pts])
    val, der = np.zeros((len(x), p + 1)), np.zeros((len(x), p + 1))
    lib.gfh_fe_basis_eval(p, len(x), x.ctypes.data, val.ctypes.data, der.ctypes.data)
    v_ref, d_ref = rf.lagrange_1d(p, x)
    assert np.abs(val - v_ref).max() < 1e-13
    assert np.abs(der - d_ref).max() < 1e-11
    assert np.abs(val.sum(axis=1) - 1).max() < 1e-13 and np.abs(der.sum(axis=1)).max() < 1e-11
    assert np.abs(val[-(p + 1):] - np.eye(p + 1)).max() < 1e-14  # Kronecker at the support points


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2), (3, 3),
                                   (3, 4)])
def test_hierarchical_and_system_numbering(dim, p):
    lib = _lib()
    npc = (p + 1) ** dim
    lex = np.zeros(npc * 3, dtype=np.int32)
    node_of = np.zeros(npc * dim, dtype=np.int32)
    comp_of = np.zeros(npc * dim, dtype=np.int32)
    assert lib.gfh_fe_numbering(dim, p, lex.ctypes.data, node_of.ctypes.data,
                                comp_of.ctypes.data) == npc
    ref = rf.hierarchical_nodes(dim, p)
    assert [tuple(t[:dim]) for t in lex.reshape(npc, 3).tolist()] == ref
    assert list(zip(node_of.tolist(), comp_of.tolist())) == rf.system_to_node_component(dim, p)
    if p <= 2:      # one scalar DoF per entity: node-major, component-minor
        assert node_of.tolist() == [i // dim for i in range(npc * dim)]


@pytest.mark.parametrize("dim,p,numbering", [(2, 3, "cellwise"), (2, 4, "component_wise"),
                                             (3, 3, "lexicographic"), (3, 3, "cellwise"),
                                             (2, 2, "cellwise")])
def test_mesh_dofs_sit_on_their_support_points(dim, p, numbering):
    """Every cell's local DoF i must be the global DoF whose support point is the cell's image of
    the unit support point of (node_of[i]) and whose component is comp_of[i]."""
    reps = [3, 2, 2][:dim]
    p0, p1 = [0.1, -0.2, 0.3][:dim], [1.3, 0.7, 0.9][:dim]
    mesh = StructuredMesh(dim, p, reps, p0, p1, numbering)
    npc, dpc = (p + 1) ** dim, dim * (p + 1) ** dim
    assert mesh.dofs_per_cell == dpc and mesh.n_dofs == dim * np.prod([r * p + 1 for r in reps])
    nodes = rf.hierarchical_nodes(dim, p)
    s2c = rf.system_to_node_component(dim, p)
    x1 = rf.support_points_1d(p)
    cd = mesh.cell_dofs.reshape(mesh.n_cells, dpc)
    cv = mesh.cell_vertices.reshape(mesh.n_cells, 1 << dim, dim)
    comp_of_global = {}
    for c in range(mesh.n_cells):
        v0, v1 = cv[c, 0], cv[c, -1]
        for i, (a, comp) in enumerate(s2c):
            x = np.array([v0[d] + x1[nodes[a][d]] * (v1[d] - v0[d]) for d in range(dim)])
            g = cd[c, i]
            assert np.abs(mesh.support_points[g] - x).max() < 1e-13
            assert comp_of_global.setdefault(g, comp) == comp
    assert sorted(comp_of_global) == list(range(mesh.n_dofs))      # every DoF is referenced
    # the dim components of a node share the support point
    sp = mesh.support_points
    order = np.lexsort(tuple(np.round(sp[:, d], 12) for d in range(dim)))
    grouped = sp[order].reshape(-1, dim, dim)
    assert np.abs(grouped - grouped[:, :1, :]).max() < 1e-13
