"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")
_lib = None


class OrcDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("degree", C.c_int32), ("model", C.c_int32),
        ("n_dofs", C.c_int64), ("n_cells", C.c_int64),
        ("cell_dofs", C.c_void_p), ("cell_vertices", C.c_void_p), ("constrained", C.c_void_p),
        ("n_iface_faces", C.c_int64), ("iface_cell", C.c_void_p), ("iface_face_no", C.c_void_p),
        ("n_iface_nodes", C.c_int64), ("iface_dofs", C.c_void_p),
        ("mu", C.c_double), ("nu", C.c_double), ("rho", C.c_double),
        ("body_force", C.c_double * 3),
        ("beta", C.c_double), ("gamma", C.c_double), ("theta", C.c_double), ("delta_t", C.c_double),
        ("type_lin", C.c_int32), ("tol_lin", C.c_double), ("max_iterations_lin", C.c_double),
        ("max_iterations_NR", C.c_int32), ("tol_f", C.c_double), ("tol_u", C.c_double),
        ("data_consistent", C.c_int32),
    ]


# vector / matrix ids (oracle.h)
NL_TOTAL_DISPLACEMENT, NL_TOTAL_DISPLACEMENT_OLD, NL_VELOCITY, NL_VELOCITY_OLD, NL_ACCELERATION, \
    NL_ACCELERATION_OLD, NL_EXTERNAL_STRESS, NL_SYSTEM_RHS, NL_SOLUTION_DELTA, NL_NEWTON_UPDATE = range(10)
LIN_OLD_VELOCITY, LIN_VELOCITY, LIN_OLD_DISPLACEMENT, LIN_DISPLACEMENT, LIN_OLD_STRESS, LIN_STRESS, \
    LIN_SYSTEM_RHS, LIN_BODY_FORCE = range(16, 24)
MAT_TANGENT, MAT_STIFFNESS, MAT_MASS, MAT_STEPPING, MAT_SYSTEM = range(5)


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(_DIR, f) for f in ("oracle.cpp", "oracle.h")]
        if (not os.path.exists(_LIB_PATH)) or any(
                os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
            subprocess.check_call(["make", "-C", _DIR], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcDesc)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_nnz.restype = C.c_int64
        L.orc_nnz.argtypes = [C.c_void_p]
        L.orc_get_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_values.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for n in ("orc_nl_update_acceleration", "orc_nl_update_velocity", "orc_nl_update_old_variables",
                  "orc_lin_assemble_system", "orc_lin_assemble_rhs", "orc_lin_update_displacement",
                  "orc_save_state", "orc_reload_state"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = None
        L.orc_nl_assemble_system.argtypes = [C.c_void_p, C.c_int]
        L.orc_nl_error_residual.restype = C.c_double
        L.orc_nl_error_residual.argtypes = [C.c_void_p]
        L.orc_nl_solve_linear_system.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.orc_nl_solve_nonlinear_timestep.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_nl_timestep.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_lin_solve.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.orc_format_precice_to_deal.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_format_deal_to_precice.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_material.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_nl_cell.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vmult.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_postprocess.restype = None
        L.orc_threads_available.restype = C.c_int
        _lib = L
    return _lib


def material(dim, mu, nu, det_F, b_bar):
    n = dim * (dim + 1) // 2
    b = np.ascontiguousarray(b_bar, dtype=np.float64)
    psi = C.c_double()
    tau = np.zeros(n)
    Jc = np.zeros((n, n))
    lib().orc_material(dim, mu, nu, det_F, b.ctypes.data, C.byref(psi), tau.ctypes.data, Jc.ctypes.data)
    return psi.value, tau, Jc


class Oracle:
    """CPU restatement of Solid / ElastoDynamics hot members + Adapter bodies for one Problem."""

    def __init__(self, problem, n_threads=None):
        L = lib()
        p = problem.params
        self.problem = problem
        self.n = problem.n_dofs
        self.n_threads = n_threads or max(1, L.orc_threads_available())
        self._keep = [np.ascontiguousarray(problem.mesh.cell_dofs, dtype=np.int32),
                      np.ascontiguousarray(problem.mesh.cell_vertices, dtype=np.float64),
                      np.ascontiguousarray(problem.constrained, dtype=np.uint8),
                      np.ascontiguousarray(problem.iface_cell, dtype=np.int32),
                      np.ascontiguousarray(problem.iface_face_no, dtype=np.int32),
                      np.ascontiguousarray(problem.iface_dofs, dtype=np.int32)]
        d = OrcDesc()
        d.dim, d.degree, d.model = problem.dim, problem.degree, problem.model
        d.n_dofs, d.n_cells = problem.n_dofs, problem.mesh.n_cells
        d.cell_dofs, d.cell_vertices, d.constrained = (a.ctypes.data for a in self._keep[:3])
        d.n_iface_faces = len(problem.iface_cell)
        d.iface_cell, d.iface_face_no = self._keep[3].ctypes.data, self._keep[4].ctypes.data
        d.n_iface_nodes = problem.n_iface_nodes
        d.iface_dofs = self._keep[5].ctypes.data
        d.mu, d.nu, d.rho = p.mu, p.nu, p.rho
        d.body_force = (C.c_double * 3)(*p.body_force)
        d.beta, d.gamma, d.theta, d.delta_t = p.beta, p.gamma, p.theta, p.delta_t
        d.type_lin = 0 if p.type_lin == "CG" else 1
        d.tol_lin, d.max_iterations_lin = p.tol_lin, p.max_iterations_lin
        d.max_iterations_NR, d.tol_f, d.tol_u = p.max_iterations_NR, p.tol_f, p.tol_u
        d.data_consistent = 1 if p.data_consistent else 0
        self._h = L.orc_create(C.byref(d))
        if not self._h:
            raise RuntimeError("oracle: " + L.orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- data access
    def pattern(self):
        nnz = lib().orc_nnz(self._h)
        rowptr = np.zeros(self.n + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int32)
        lib().orc_get_pattern(self._h, rowptr.ctypes.data, col.ctypes.data)
        return rowptr, col

    def values(self, which):
        val = np.zeros(lib().orc_nnz(self._h))
        lib().orc_get_values(self._h, which, val.ctypes.data)
        return val

    def csr(self, which):
        import scipy.sparse as sp
        rowptr, col = self.pattern()
        return sp.csr_matrix((self.values(which), col, rowptr), shape=(self.n, self.n))

    def get(self, which):
        out = np.zeros(self.n)
        lib().orc_get_vector(self._h, which, out.ctypes.data)
        return out

    def set(self, which, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.n,)
        lib().orc_set_vector(self._h, which, v.ctypes.data)

    # -- nonlinear
    def nl_update_acceleration(self): lib().orc_nl_update_acceleration(self._h)
    def nl_update_velocity(self): lib().orc_nl_update_velocity(self._h)
    def nl_update_old_variables(self): lib().orc_nl_update_old_variables(self._h)
    def nl_assemble_system(self): lib().orc_nl_assemble_system(self._h, self.n_threads)
    def nl_error_residual(self): return lib().orc_nl_error_residual(self._h)

    def nl_solve_linear_system(self):
        it, res = C.c_uint32(), C.c_double()
        st = lib().orc_nl_solve_linear_system(self._h, C.byref(it), C.byref(res))
        return st, it.value, res.value

    def nl_timestep(self):
        """Body of the coupling loop (nonlinear_elasticity.cc:121,138-144). Returns
        (n_newton_solves, history rows)."""
        hist = np.zeros((32, 6))
        n = lib().orc_nl_timestep(self._h, self.n_threads, hist.ctypes.data, 32)
        if n < 0:
            raise RuntimeError({-1: "No convergence in nonlinear solver!", -2: "CG did not converge",
                                -3: lib().orc_last_error().decode()}[n])
        return n, hist[:n]

    def nl_cell(self, cell, u_local, acc_local):
        dpc = self.problem.mesh.dofs_per_cell
        u = np.ascontiguousarray(u_local, dtype=np.float64)
        a = np.ascontiguousarray(acc_local, dtype=np.float64)
        K = np.zeros((dpc, dpc))
        r = np.zeros(dpc)
        lib().orc_nl_cell(self._h, cell, u.ctypes.data, a.ctypes.data, K.ctypes.data, r.ctypes.data)
        return K, r

    # -- linear
    def lin_assemble_system(self): lib().orc_lin_assemble_system(self._h)
    def lin_assemble_rhs(self): lib().orc_lin_assemble_rhs(self._h)
    def lin_update_displacement(self): lib().orc_lin_update_displacement(self._h)

    def lin_solve(self):
        it, res = C.c_uint32(), C.c_double()
        st = lib().orc_lin_solve(self._h, C.byref(it), C.byref(res))
        return st, it.value, res.value

    def lin_step(self):
        """assemble_rhs + solve + update_displacement (linear_elasticity.cc:680-686)."""
        self.lin_assemble_rhs()
        st, it, res = self.lin_solve()
        if st != 0:
            raise RuntimeError("CG did not converge")
        self.lin_update_displacement()
        return it, res

    # -- adapter bodies
    def format_precice_to_deal(self, buf, which):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        lib().orc_format_precice_to_deal(self._h, buf.ctypes.data, which)

    def format_deal_to_precice(self, which):
        out = np.zeros(self.problem.n_iface_nodes * self.problem.dim)
        lib().orc_format_deal_to_precice(self._h, which, out.ctypes.data)
        return out

    def save_state(self): lib().orc_save_state(self._h)
    def reload_state(self): lib().orc_reload_state(self._h)

    def postprocess(self, which):
        """output_results + Postprocessor: (points [n_cells, npts, dim] = X + u on the displaced
        grid, fields [n_cells, npts, dim + dim*dim] = u | strain(d*dim+e))."""
        dim = self.problem.dim
        npts = (self.problem.params.poly_degree + 1) ** dim
        nc = self.problem.mesh.n_cells
        pts = np.zeros((nc, npts, dim))
        fld = np.zeros((nc, npts, dim + dim * dim))
        lib().orc_postprocess(self._h, which, pts.ctypes.data, fld.ctypes.data)
        return pts, fld

    def vmult(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n)
        lib().orc_vmult(self._h, which, x.ctypes.data, y.ctypes.data)
        return y
