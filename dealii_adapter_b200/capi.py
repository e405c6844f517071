"""ctypes binding of the C-ABI in include/graft_fem.h (libgraftfem.so, sm_100a CUDA).

There is no CPU fallback: loading fails loudly when the library is missing, and every call
raises GraftError with the library's message on a non-zero status.
"""
import ctypes as C
import os

import numpy as np

from . import build

GF_OK = 0
GF_ERR_INVALID_ARG, GF_ERR_CUDA, GF_ERR_NOT_CONVERGED, GF_ERR_DET_F, GF_ERR_NCCL, GF_ERR_UNSUPPORTED = range(1, 7)

# vector ids
NL_TOTAL_DISPLACEMENT, NL_TOTAL_DISPLACEMENT_OLD, NL_VELOCITY, NL_VELOCITY_OLD, NL_ACCELERATION, \
    NL_ACCELERATION_OLD, NL_EXTERNAL_STRESS, NL_SYSTEM_RHS, NL_SOLUTION_DELTA, NL_NEWTON_UPDATE = range(10)
LIN_OLD_VELOCITY, LIN_VELOCITY, LIN_OLD_DISPLACEMENT, LIN_DISPLACEMENT, LIN_OLD_STRESS, LIN_STRESS, \
    LIN_SYSTEM_RHS, LIN_BODY_FORCE = range(16, 24)
VEC_SCRATCH0, VEC_SCRATCH1 = 28, 29
MAT_TANGENT, MAT_STIFFNESS, MAT_MASS, MAT_SYSTEM, MAT_MG_F32 = range(5)
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_JACOBI, PRECOND_MULTIGRID = range(4)
OPT_PRECONDITIONER, OPT_CG_CHECK_INTERVAL, OPT_PROFILE, OPT_OPERATOR, OPT_SPMV_KERNEL, \
    OPT_MG_SMOOTHER_DEGREE, OPT_MG_COARSE_DEGREE, OPT_MG_SMOOTHER_RATIO, \
    OPT_CG_INITIAL_GUESS, OPT_MG_MATRIX_PRECISION, OPT_DIRECT_SOLVER = range(11)

EXPORTED_SYMBOLS = [
    "gf_create", "gf_destroy", "gf_last_error", "gf_set_option", "gf_comm_unique_id",
    "gf_comm_create", "gf_comm_destroy", "gf_set_traction", "gf_get_interface_displacement",
    "gf_state_save", "gf_state_restore", "gf_nl_begin_step", "gf_nl_newton_assemble",
    "gf_nl_newton_solve", "gf_nl_end_step", "gf_lin_assemble_once", "gf_lin_step", "gf_get_vector",
    "gf_set_vector", "gf_nnz", "gf_export_csr", "gf_spmv", "gf_spmv_timed", "gf_profile_get",
    "gf_synchronize", "gf_event_record", "gf_event_elapsed_ms", "gf_mg_attach", "gf_mg_vcycle",
    "gf_comm_transport", "gf_comm_timed", "gf_postprocess", "gf_export_rows", "gf_comm_ipc_begin",
    "gf_comm_ipc_finish", "gf_direct_info",
]


class GraftError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("graft_fem error %d: %s" % (code, msg))
        self.code = code


class GfDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("degree", C.c_int32), ("model", C.c_int32),
        ("n_dofs", C.c_int64), ("n_cells", C.c_int64),
        ("cell_dofs", C.c_void_p), ("cell_vertices", C.c_void_p), ("constrained", C.c_void_p),
        ("n_iface_faces", C.c_int64), ("iface_cell", C.c_void_p), ("iface_face_no", C.c_void_p),
        ("n_iface_nodes", C.c_int64), ("iface_dofs", C.c_void_p),
        ("mu", C.c_double), ("nu", C.c_double), ("rho", C.c_double),
        ("body_force", C.c_double * 3),
        ("beta", C.c_double), ("gamma", C.c_double), ("theta", C.c_double), ("delta_t", C.c_double),
        ("data_consistent", C.c_int32), ("device", C.c_int32),
        ("n_owned_dofs", C.c_int64), ("comm", C.c_void_p),
        ("n_neighbors", C.c_int32), ("nbr_rank", C.c_void_p),
        ("send_ptr", C.c_void_p), ("send_dofs", C.c_void_p),
        ("recv_ptr", C.c_void_p), ("recv_dofs", C.c_void_p),
        ("dof_global", C.c_void_p), ("slab_axis", C.c_int32),
        ("n_constraint_lines", C.c_int64), ("line_dof", C.c_void_p), ("line_ptr", C.c_void_p),
        ("line_master", C.c_void_p), ("line_weight", C.c_void_p),
    ]


class GfProfile(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("assemble_cells_ms", "assemble_faces_ms", "scatter_ms", "spmv_ms", "cg_vector_ms",
                 "update_ms", "halo_ms")] + \
               [(n, C.c_int64) for n in
                ("assemble_cells_launches", "assemble_faces_launches", "scatter_launches",
                 "spmv_launches", "cg_vector_launches", "update_launches", "halo_launches",
                 "kernel_launches")] + \
               [("mg_spmv_ms", C.c_double), ("mg_vector_ms", C.c_double),
                ("mg_spmv_launches", C.c_int64), ("mg_vector_launches", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    """Load libgraftfem.so (built in-tree by build.build_cuda()). No fallback."""
    global _lib
    if _lib is None:
        path = build.LIB_CUDA
        if not os.path.exists(path):
            raise ImportError(
                "libgraftfem.so is missing (%s). Build it with `python -m dealii_adapter_b200.build` "
                "or __graft_entry__.build(); there is no CPU fallback." % path)
        L = C.CDLL(path)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.gf_create.argtypes = [C.POINTER(GfDesc), C.POINTER(vp)]
        L.gf_destroy.argtypes = [vp]
        L.gf_destroy.restype = None
        L.gf_last_error.argtypes = [vp]
        L.gf_last_error.restype = C.c_char_p
        L.gf_set_option.argtypes = [vp, i32, i64]
        L.gf_comm_unique_id.argtypes = [vp]
        L.gf_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
        L.gf_comm_destroy.argtypes = [vp]
        L.gf_comm_destroy.restype = None
        L.gf_comm_ipc_begin.argtypes = [i32, i32, i32, C.POINTER(vp), vp]
        L.gf_comm_ipc_finish.argtypes = [vp, vp, i32]
        L.gf_set_traction.argtypes = [vp, vp]
        L.gf_get_interface_displacement.argtypes = [vp, vp]
        for n in ("gf_state_save", "gf_state_restore", "gf_nl_begin_step", "gf_nl_end_step",
                  "gf_lin_assemble_once", "gf_synchronize"):
            getattr(L, n).argtypes = [vp]
        L.gf_nl_newton_assemble.argtypes = [vp, C.POINTER(dbl)]
        L.gf_nl_newton_solve.argtypes = [vp, i32, dbl, dbl, C.POINTER(C.c_uint32), C.POINTER(dbl),
                                         C.POINTER(dbl)]
        L.gf_lin_step.argtypes = [vp, i32, dbl, C.POINTER(C.c_uint32), C.POINTER(dbl)]
        L.gf_get_vector.argtypes = [vp, i32, vp]
        L.gf_set_vector.argtypes = [vp, i32, vp]
        L.gf_nnz.argtypes = [vp]
        L.gf_nnz.restype = i64
        L.gf_export_csr.argtypes = [vp, i32, vp, vp, vp]
        L.gf_export_rows.argtypes = [vp, i32, i64, vp, vp, vp, vp]
        L.gf_spmv.argtypes = [vp, i32, i32, i32]
        L.gf_spmv_timed.argtypes = [vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
        L.gf_profile_get.argtypes = [vp, C.POINTER(GfProfile), i32]
        L.gf_event_record.argtypes = [vp, i32]
        L.gf_event_elapsed_ms.argtypes = [vp, i32, i32, C.POINTER(dbl)]
        L.gf_mg_attach.argtypes = [vp, vp, vp]
        L.gf_mg_vcycle.argtypes = [vp, i32, i32]
        L.gf_postprocess.argtypes = [vp, i32, vp]
        _lib = L
    return _lib


class Comm:
    """NCCL communicator handle for the partitioned path; the 128-byte id is broadcast by the host
    (torch.distributed in bench.py / tests)."""

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int):
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        rc = lib().gf_comm_create(buf, rank, n_ranks, device, C.byref(h))
        if rc != GF_OK:
            raise GraftError(rc, "gf_comm_create failed")
        self._h = h
        self.rank, self.n_ranks = rank, n_ranks

    @classmethod
    def from_ipc(cls, rank: int, n_ranks: int, device: int, all_gather, share_device=False):
        """Communicator without NCCL (gf_comm_ipc_begin / _finish): `all_gather(bytes) -> [bytes]`
        moves every rank's 64-byte window handle over the host's own channel (e.g.
        torch.distributed with the gloo backend). Ranks may share a device."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        mine = (C.c_uint8 * 64)()
        rc = lib().gf_comm_ipc_begin(rank, n_ranks, device, C.byref(h), mine)
        if rc != GF_OK:
            raise GraftError(rc, "gf_comm_ipc_begin failed")
        handles = all_gather(bytes(mine))
        assert len(handles) == n_ranks and all(len(x) == 64 for x in handles)
        buf = (C.c_uint8 * (64 * n_ranks)).from_buffer_copy(b"".join(handles))
        rc = lib().gf_comm_ipc_finish(h, buf, 1 if share_device else 0)
        if rc != GF_OK:
            raise GraftError(rc, "gf_comm_ipc_finish failed (cudaIpcOpenMemHandle)")
        self._h = h
        self.rank, self.n_ranks = rank, n_ranks
        return self

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().gf_comm_unique_id(buf)
        if rc != GF_OK:
            raise GraftError(rc, "gf_comm_unique_id failed")
        return bytes(buf)

    def transport(self):
        """("peer_windows" | "nccl", halo exchanges issued, all-reduces issued)."""
        nh, na = C.c_int64(0), C.c_int64(0)
        f = lib().gf_comm_transport
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        f.restype = C.c_int
        t = f(self._h, C.byref(nh), C.byref(na))
        return ("peer_windows" if t == 1 else "nccl"), nh.value, na.value

    def close(self):
        if self._h:
            lib().gf_comm_destroy(self._h)
            self._h = None


class Handle:
    """Owns one gf_handle. `problem` is a dealii_adapter_b200.problem.Problem; `partition` an
    optional mesh.MeshPartition with `comm`. `slab_axis` (0, 1, 2 or None): the axis the mesh is
    (or would be) cut along - with it every dot product / norm / restriction sums over a
    partition-independent tree, so a run gives the same bits for any number of ranks (gf_desc)."""

    def __init__(self, problem, device=0, partition=None, comm=None, slab_axis=None):
        L = lib()
        p = problem.params
        self.problem = problem
        self.partition = partition
        d = GfDesc()
        d.dim, d.degree, d.model = problem.dim, problem.degree, problem.model
        if partition is None:
            cell_dofs = problem.mesh.cell_dofs
            cell_vertices = problem.mesh.cell_vertices
            constrained = problem.constrained
            iface_cell, iface_face_no = problem.iface_cell, problem.iface_face_no
            iface_dofs = problem.iface_dofs
            n_dofs = n_owned = problem.n_dofs
            n_cells = problem.mesh.n_cells
            self.local_to_global = None
        else:
            l2g = partition.local_to_global
            g2l = -np.ones(problem.n_dofs, dtype=np.int64)
            g2l[l2g] = np.arange(len(l2g))
            cell_dofs, cell_vertices = partition.cell_dofs, partition.cell_vertices
            constrained = problem.constrained[l2g]
            gc2l = -np.ones(problem.mesh.n_cells, dtype=np.int64)
            gc2l[partition.local_cell_global] = np.arange(partition.n_local_cells)
            keep = gc2l[problem.iface_cell] >= 0
            iface_cell = gc2l[problem.iface_cell][keep].astype(np.int32)
            iface_face_no = problem.iface_face_no[keep]
            # interface nodes visible on this rank, in the caller's (global) IndexSet order
            vis = g2l[problem.iface_dofs[0]] >= 0
            iface_dofs = g2l[problem.iface_dofs[:, vis]].astype(np.int32)
            self.iface_visible = vis
            n_dofs, n_owned = partition.n_local_dofs, partition.n_owned_dofs
            n_cells = partition.n_local_cells
            self.local_to_global = l2g
        self.n_dofs, self.n_owned = int(n_dofs), int(n_owned)
        self.n_iface_nodes = int(iface_dofs.shape[1])
        keep_alive = [np.ascontiguousarray(cell_dofs, dtype=np.int32),
                      np.ascontiguousarray(cell_vertices, dtype=np.float64),
                      np.ascontiguousarray(constrained, dtype=np.uint8),
                      np.ascontiguousarray(iface_cell, dtype=np.int32),
                      np.ascontiguousarray(iface_face_no, dtype=np.int32),
                      np.ascontiguousarray(iface_dofs, dtype=np.int32)]
        d.n_dofs, d.n_cells = n_dofs, n_cells
        self.n_cells = n_cells
        d.cell_dofs, d.cell_vertices, d.constrained = (a.ctypes.data for a in keep_alive[:3])
        d.n_iface_faces = len(keep_alive[3])
        d.iface_cell, d.iface_face_no = keep_alive[3].ctypes.data, keep_alive[4].ctypes.data
        d.n_iface_nodes = self.n_iface_nodes
        d.iface_dofs = keep_alive[5].ctypes.data
        d.mu, d.nu, d.rho = p.mu, p.nu, p.rho
        d.body_force = (C.c_double * 3)(*p.body_force)
        d.beta, d.gamma, d.theta, d.delta_t = p.beta, p.gamma, p.theta, p.delta_t
        d.data_consistent = 1 if p.data_consistent else 0
        d.device = device
        d.n_owned_dofs = n_owned
        if slab_axis is None and partition is not None:
            slab_axis = partition.axis
        d.slab_axis = 0 if slab_axis is None else slab_axis + 1
        if partition is not None:
            keep_alive.append(np.ascontiguousarray(partition.local_to_global, dtype=np.int64))
            d.dof_global = keep_alive[-1].ctypes.data
        if partition is not None and comm is not None:
            d.comm = comm._h
            d.n_neighbors = len(partition.nbr_rank)
            keep_alive += [np.ascontiguousarray(partition.nbr_rank, dtype=np.int32),
                           np.ascontiguousarray(partition.send_ptr, dtype=np.int64),
                           np.ascontiguousarray(partition.send_dofs, dtype=np.int32),
                           np.ascontiguousarray(partition.recv_ptr, dtype=np.int64),
                           np.ascontiguousarray(partition.recv_dofs, dtype=np.int32)]
            d.nbr_rank, d.send_ptr, d.send_dofs, d.recv_ptr, d.recv_dofs = (
                a.ctypes.data for a in keep_alive[-5:])
        # hanging-node constraint lines (problem.extra["constraint_lines"] = (dof, ptr, master,
        # weight) in the caller's numbering; the structured stand-in meshes have none)
        lines = getattr(problem, "extra", {}).get("constraint_lines")
        if lines is not None:
            if partition is not None:
                raise ValueError("hanging-node constraints need a serial handle")
            keep_alive += [np.ascontiguousarray(lines[0], dtype=np.int32),
                           np.ascontiguousarray(lines[1], dtype=np.int64),
                           np.ascontiguousarray(lines[2], dtype=np.int32),
                           np.ascontiguousarray(lines[3], dtype=np.float64)]
            d.n_constraint_lines = len(keep_alive[-4])
            d.line_dof, d.line_ptr, d.line_master, d.line_weight = (
                a.ctypes.data for a in keep_alive[-4:])
        h = C.c_void_p()
        rc = L.gf_create(C.byref(d), C.byref(h))
        if rc != GF_OK:
            raise GraftError(rc, L.gf_last_error(None).decode())
        self._h = h

    # -- plumbing
    def _check(self, rc):
        if rc != GF_OK:
            raise GraftError(rc, lib().gf_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().gf_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_option(self, opt, value):
        self._check(lib().gf_set_option(self._h, opt, int(value)))

    def mg_attach(self, coarse, child_cells):
        """Link `coarse` (a Handle of the next coarser refinement level) below this level;
        child_cells[coarse_cell, k] = this level's local cell index of cell->child(k) or -1."""
        cc = np.ascontiguousarray(child_cells, dtype=np.int32)
        self._check(lib().gf_mg_attach(self._h, coarse._h, cc.ctypes.data))
        self._mg_coarse = coarse  # keep the coarse handle alive as long as this one

    def mg_vcycle(self, which_b, which_x):
        self._check(lib().gf_mg_vcycle(self._h, which_b, which_x))

    # -- Adapter bodies
    def set_traction(self, iface_buf):
        buf = np.ascontiguousarray(iface_buf, dtype=np.float64)
        assert buf.size == self.n_iface_nodes * self.problem.dim
        self._check(lib().gf_set_traction(self._h, buf.ctypes.data))

    def get_interface_displacement(self):
        out = np.zeros(self.n_iface_nodes * self.problem.dim)
        self._check(lib().gf_get_interface_displacement(self._h, out.ctypes.data))
        return out

    def state_save(self):
        self._check(lib().gf_state_save(self._h))

    def state_restore(self):
        self._check(lib().gf_state_restore(self._h))

    # -- Solid
    def nl_begin_step(self):
        self._check(lib().gf_nl_begin_step(self._h))

    def nl_newton_assemble(self):
        r = C.c_double()
        self._check(lib().gf_nl_newton_assemble(self._h, C.byref(r)))
        return r.value

    def nl_newton_solve(self, type_lin, tol_lin, max_iterations_lin):
        it, res, upd = C.c_uint32(), C.c_double(), C.c_double()
        self._check(lib().gf_nl_newton_solve(self._h, type_lin, tol_lin, max_iterations_lin,
                                             C.byref(it), C.byref(res), C.byref(upd)))
        return it.value, res.value, upd.value

    def nl_end_step(self):
        self._check(lib().gf_nl_end_step(self._h))

    # -- ElastoDynamics
    def lin_assemble_once(self):
        self._check(lib().gf_lin_assemble_once(self._h))

    def lin_step(self, type_lin, max_iterations_lin):
        it, res = C.c_uint32(), C.c_double()
        self._check(lib().gf_lin_step(self._h, type_lin, max_iterations_lin, C.byref(it),
                                      C.byref(res)))
        return it.value, res.value

    # -- data access
    def get_vector(self, which):
        out = np.zeros(self.n_dofs)
        self._check(lib().gf_get_vector(self._h, which, out.ctypes.data))
        return out

    def set_vector(self, which, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.n_dofs,)
        self._check(lib().gf_set_vector(self._h, which, v.ctypes.data))

    def nnz(self):
        return lib().gf_nnz(self._h)

    def export_csr(self, which):
        nnz = self.nnz()
        rowptr = np.zeros(self.n_owned + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        self._check(lib().gf_export_csr(self._h, which, rowptr.ctypes.data, col.ctypes.data,
                                        val.ctypes.data))
        return rowptr, col, val

    def export_rows(self, which, rows):
        """Selected owned rows (caller dof ids) of one matrix: (rowptr, col, val), ascending
        caller columns per row."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
        self._check(lib().gf_export_rows(self._h, which, len(rows), rows.ctypes.data,
                                         rowptr.ctypes.data, None, None))
        col = np.zeros(rowptr[-1], dtype=np.int32)
        val = np.zeros(rowptr[-1])
        self._check(lib().gf_export_rows(self._h, which, len(rows), rows.ctypes.data,
                                         rowptr.ctypes.data, col.ctypes.data, val.ctypes.data))
        return rowptr, col, val

    def postprocess(self, which_vector):
        """Output-step fields of DataOut + Postprocessor (postprocessor.h:44-76):
        [n_cells, (p+1)^dim patch points, dim + dim*dim] = displacement, then strain (d*dim+e)."""
        dim = self.problem.dim
        npts = (self.problem.params.poly_degree + 1) ** dim
        out = np.zeros((self.n_cells, npts, dim + dim * dim))
        self._check(lib().gf_postprocess(self._h, which_vector, out.ctypes.data))
        return out

    def spmv(self, which_matrix, which_x, which_y):
        self._check(lib().gf_spmv(self._h, which_matrix, which_x, which_y))

    def spmv_timed(self, which_matrix, n_reps):
        ms, nbytes = C.c_double(), C.c_double()
        self._check(lib().gf_spmv_timed(self._h, which_matrix, n_reps, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def comm_timed(self, n_reps):
        """(halo exchange us, 2-scalar all-reduce us) averaged over n_reps; collective."""
        a, b = C.c_double(), C.c_double()
        f = lib().gf_comm_timed
        f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        f.restype = C.c_int
        self._check(f(self._h, n_reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def direct_info(self):
        """(solves answered by the band Cholesky, half bandwidth, last relative residual);
        GraftError(GF_ERR_UNSUPPORTED) with the reason when the CG stand-in runs instead."""
        n, w, r = C.c_int64(), C.c_int64(), C.c_double()
        f = lib().gf_direct_info
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        f.restype = C.c_int
        self._check(f(self._h, C.byref(n), C.byref(w), C.byref(r)))
        return n.value, w.value, r.value

    def profile(self, reset=False):
        p = GfProfile()
        self._check(lib().gf_profile_get(self._h, C.byref(p), 1 if reset else 0))
        return p.as_dict()

    def synchronize(self):
        self._check(lib().gf_synchronize(self._h))

    def event_record(self, slot):
        self._check(lib().gf_event_record(self._h, slot))

    def event_elapsed_ms(self, slot_begin, slot_end):
        ms = C.c_double()
        self._check(lib().gf_event_elapsed_ms(self._h, slot_begin, slot_end, C.byref(ms)))
        return ms.value
