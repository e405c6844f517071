"""ctypes view of the C++ structured-mesh host scaffolding (host/structured_mesh.{h,cc}).

It stands in for deal.II's Triangulation/DoFHandler setup (reference: make_grid/system_setup,
nonlinear_elasticity.cc:171-380, linear_elasticity.cc:79-244) and produces exactly the arrays the
C-ABI (include/graft_fem.h) and the CPU oracle consume.
"""
import ctypes as C
import os

import numpy as np

from . import build

_lib = None


def host_lib():
    global _lib
    if _lib is None:
        path = build.LIB_HOST
        if not os.path.exists(path):
            build.build_host()
        lib = C.CDLL(path)
        lib.gfh_mesh_create.restype = C.c_void_p
        lib.gfh_mesh_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.c_int]
        lib.gfh_mesh_destroy.argtypes = [C.c_void_p]
        for name in ("n_cells", "n_dofs", "n_nodes"):
            f = getattr(lib, "gfh_mesh_" + name)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        lib.gfh_mesh_dofs_per_cell.restype = C.c_int
        lib.gfh_mesh_dofs_per_cell.argtypes = [C.c_void_p]
        lib.gfh_mesh_cell_dofs.restype = C.POINTER(C.c_int32)
        lib.gfh_mesh_cell_dofs.argtypes = [C.c_void_p]
        lib.gfh_mesh_cell_vertices.restype = C.POINTER(C.c_double)
        lib.gfh_mesh_cell_vertices.argtypes = [C.c_void_p]
        lib.gfh_mesh_support_points.restype = C.POINTER(C.c_double)
        lib.gfh_mesh_support_points.argtypes = [C.c_void_p]
        lib.gfh_mesh_boundary_dof_mask.restype = C.c_int64
        lib.gfh_mesh_boundary_dof_mask.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p]
        lib.gfh_mesh_boundary_faces.restype = C.c_int64
        lib.gfh_mesh_boundary_faces.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        lib.gfh_mesh_interface_dofs.restype = C.c_int64
        lib.gfh_mesh_interface_dofs.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        lib.gfh_partition_create.restype = C.c_void_p
        lib.gfh_partition_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.gfh_partition_destroy.argtypes = [C.c_void_p]
        lib.gfh_partition_size.restype = C.c_int64
        lib.gfh_partition_size.argtypes = [C.c_void_p, C.c_int]
        lib.gfh_partition_array.restype = C.c_void_p
        lib.gfh_partition_array.argtypes = [C.c_void_p, C.c_int]
        lib.gfh_write_vtk.restype = C.c_int
        lib.gfh_write_vtk.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
        lib.gfh_vtk_point_index_from_ijk.restype = C.c_int
        lib.gfh_vtk_point_index_from_ijk.argtypes = [C.c_int] * 6
        _lib = lib
    return _lib


NUMBERING = {"cellwise": 0, "component_wise": 1, "lexicographic": 2}

# colorize boundary ids (GridGenerator::subdivided_hyper_rectangle): bit f = face id f
XM, XP, YM, YP, ZM, ZP = (1 << 0), (1 << 1), (1 << 2), (1 << 3), (1 << 4), (1 << 5)


def scenario_geometry(scenario, dim, flap_location=0.0):
    """Box corners and boundary roles of the two shipped scenarios
    (nonlinear_elasticity.cc:189-226,269-278; linear_elasticity.cc:94-131,176-186)."""
    if scenario == "FSI3":
        p0 = [0.24899, 0.19, -0.005][:dim]
        p1 = [0.6, 0.21, 0.005][:dim]
        base_reps = [18, 3, 1][:dim]
        clamped = XM            # id_flap_short_bottom = 0
        interface = YM | YP | XP  # long bottom/top = 2,3 ; short top = 1
    elif scenario == "PF":
        p0 = [flap_location - 0.05, 0.0, 0.0][:dim]
        p1 = [flap_location + 0.05, 1.0, 0.3][:dim]
        base_reps = [3, 18, 1][:dim]
        clamped = YM            # id_flap_short_bottom = 2
        interface = XM | XP | YP
    else:
        raise ValueError("Scenario must be FSI3 or PF (parameters.cc:136-139)")
    zclamp = (ZM | ZP) if dim == 3 else 0
    return p0, p1, base_reps, clamped, interface, zclamp


class StructuredMesh:
    def __init__(self, dim, degree, reps, p0, p1, numbering="cellwise"):
        lib = host_lib()
        self.dim, self.degree = dim, degree
        self.reps = list(reps)
        self.p0, self.p1 = list(p0), list(p1)
        self.numbering = numbering
        r = (C.c_int * 3)(*(list(reps) + [1] * (3 - dim)))
        a = (C.c_double * 3)(*(list(p0) + [0.0] * (3 - dim)))
        b = (C.c_double * 3)(*(list(p1) + [0.0] * (3 - dim)))
        self._h = lib.gfh_mesh_create(dim, degree, r, a, b, NUMBERING[numbering])
        if not self._h:
            raise ValueError("invalid mesh description")
        self.n_cells = lib.gfh_mesh_n_cells(self._h)
        self.n_dofs = lib.gfh_mesh_n_dofs(self._h)
        self.n_nodes = lib.gfh_mesh_n_nodes(self._h)
        self.dofs_per_cell = lib.gfh_mesh_dofs_per_cell(self._h)
        nv = 1 << dim
        self.cell_dofs = np.ctypeslib.as_array(
            lib.gfh_mesh_cell_dofs(self._h), shape=(self.n_cells * self.dofs_per_cell,)).copy()
        self.cell_vertices = np.ctypeslib.as_array(
            lib.gfh_mesh_cell_vertices(self._h), shape=(self.n_cells * nv * dim,)).copy()
        self.support_points = np.ctypeslib.as_array(
            lib.gfh_mesh_support_points(self._h), shape=(self.n_dofs, dim)).copy()

    def __del__(self):
        if getattr(self, "_h", None):
            host_lib().gfh_mesh_destroy(self._h)
            self._h = None

    def cells_at_boundary(self):
        """uint8 [n_cells]: 1 for cells with a face on the boundary (DataOut::curved_boundary)."""
        r = list(self.reps) + [1] * (3 - self.dim)
        k, j, i = np.meshgrid(np.arange(r[2]), np.arange(r[1]), np.arange(r[0]), indexing="ij")
        at = (i == 0) | (i == r[0] - 1) | (j == 0) | (j == r[1] - 1)
        if self.dim == 3:
            at |= (k == 0) | (k == r[2] - 1)
        return np.ascontiguousarray(at.reshape(-1), dtype=np.uint8)

    def write_vtk(self, filename, fields, curved_boundary=True):
        """solution-*.vtk as output_results writes it (nonlinear_elasticity.cc:1215-1254): `fields`
        is the [n_cells, npts, dim + dim*dim] array of gf_postprocess; the C++ host writer
        (host/vtk_output.cc) builds the displaced patch points and the file."""
        f = np.ascontiguousarray(fields, dtype=np.float64)
        v = np.ascontiguousarray(self.cell_vertices, dtype=np.float64)
        at = self.cells_at_boundary() if curved_boundary else None
        rc = host_lib().gfh_write_vtk(str(filename).encode(), self.dim, self.degree, self.n_cells,
                                      v.ctypes.data, f.ctypes.data,
                                      at.ctypes.data if at is not None else None)
        if rc != 0:
            raise IOError("cannot write %s" % filename)

    def boundary_dof_mask(self, face_mask, comp_mask, mask=None):
        if mask is None:
            mask = np.zeros(self.n_dofs, dtype=np.uint8)
        host_lib().gfh_mesh_boundary_dof_mask(self._h, face_mask, comp_mask, mask.ctypes.data)
        return mask

    def boundary_faces(self, face_mask):
        lib = host_lib()
        n = lib.gfh_mesh_boundary_faces(self._h, face_mask, None, None)
        cells = np.zeros(n, dtype=np.int32)
        faces = np.zeros(n, dtype=np.int32)
        if n:
            lib.gfh_mesh_boundary_faces(self._h, face_mask, cells.ctypes.data, faces.ctypes.data)
        return cells, faces

    def interface_dofs(self, face_mask):
        lib = host_lib()
        n = lib.gfh_mesh_interface_dofs(self._h, face_mask, None)
        out = np.zeros(n * self.dim, dtype=np.int32)
        if n:
            lib.gfh_mesh_interface_dofs(self._h, face_mask, out.ctypes.data)
        return out.reshape(self.dim, n)

    def partition(self, axis, nparts, rank):
        return MeshPartition(self, axis, nparts, rank)


class MeshPartition:
    """1-D slab partition: local cells (owned + one ghost layer), local dof numbering
    [owned | ghost], neighbour exchange lists."""

    def __init__(self, mesh, axis, nparts, rank):
        lib = host_lib()
        h = lib.gfh_partition_create(mesh._h, axis, nparts, rank)
        if not h:
            raise ValueError("invalid partition request")
        try:
            size = lambda w: lib.gfh_partition_size(h, w)

            def arr(which, dtype, n):
                if n == 0:
                    return np.zeros(0, dtype=dtype)
                ptr = C.cast(lib.gfh_partition_array(h, which), C.POINTER(
                    {np.int32: C.c_int32, np.int64: C.c_int64, np.float64: C.c_double}[dtype]))
                return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

            self.rank, self.nparts, self.axis = rank, nparts, axis
            self.n_local_dofs, self.n_owned_dofs = size(0), size(1)
            self.n_local_cells = size(2)
            nn = size(3)
            nv = 1 << mesh.dim
            self.cell_dofs = arr(0, np.int32, self.n_local_cells * mesh.dofs_per_cell)
            self.cell_vertices = arr(1, np.float64, self.n_local_cells * nv * mesh.dim)
            self.local_cell_global = arr(2, np.int64, self.n_local_cells)
            self.local_to_global = arr(3, np.int32, self.n_local_dofs)
            self.nbr_rank = arr(4, np.int32, nn)
            self.send_ptr = arr(5, np.int64, nn + 1)
            self.recv_ptr = arr(6, np.int64, nn + 1)
            self.send_dofs = arr(7, np.int32, size(4))
            self.recv_dofs = arr(8, np.int32, size(5))
        finally:
            lib.gfh_partition_destroy(h)
