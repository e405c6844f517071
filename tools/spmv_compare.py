import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dealii_adapter_b200 import capi
from dealii_adapter_b200.problem import SolverParameters, make_problem
p = SolverParameters(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF", delta_t=0.01,
                     mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6, max_iterations_lin=1.0)
prob = make_problem(p, 3, reps=[24, 144, 24], numbering="lexicographic")
h = capi.Handle(prob)
h.set_traction(np.tile([2000.0, 0.0, 0.0], prob.n_iface_nodes))
h.nl_newton_assemble()
x = np.random.RandomState(0).uniform(-1, 1, prob.n_dofs)
h.set_vector(capi.VEC_SCRATCH0, x)
ys = []
for kind in (1, 0):
    h.set_option(capi.OPT_SPMV_KERNEL, kind)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    ys.append(h.get_vector(capi.VEC_SCRATCH1))
    for rep in range(3):
        ms, nb = h.spmv_timed(capi.MAT_TANGENT, 50)
        print("kernel kind %d: %.4f ms  %.1f GB/s" % (kind, ms, nb / ms / 1e6))
print("max diff LDG vs TMA:", np.abs(ys[0] - ys[1]).max(), "rel", np.abs(ys[0] - ys[1]).max() / np.abs(ys[0]).max())
h.nl_begin_step(); h.nl_newton_assemble()
t0 = time.time(); it, res, upd = h.nl_newton_solve(0, 1e-6, 1.0); print("CG TMA: its", it, "wall", time.time() - t0)
