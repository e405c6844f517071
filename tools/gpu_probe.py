"""Exploratory GPU probe: timings of the individual phases at a given mesh size (not the bench)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dealii_adapter_b200 import capi, solvers
from dealii_adapter_b200.problem import SolverParameters, make_problem

reps = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "24,144,24").split(",")]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = SolverParameters(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF", delta_t=0.01,
                     mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6, max_iterations_lin=1.0)
t0 = time.time()
prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
print("mesh: %d cells %d dofs, %.1fs" % (prob.mesh.n_cells, prob.n_dofs, time.time() - t0))
t0 = time.time()
h = capi.Handle(prob)
print("gf_create %.2fs, nnz %d" % (time.time() - t0, h.nnz()))
h.set_option(capi.OPT_PROFILE, 1)
n = prob.n_iface_nodes
h.set_traction(np.tile([2000.0, 0.0, 0.0], n))
for rep in range(3):
    h.profile(reset=True)
    t0 = time.time()
    r = h.nl_newton_assemble()
    h.synchronize()
    print("assemble wall %.2f ms, res %.6e" % ((time.time() - t0) * 1e3, r), h.profile())
ms, nbytes = h.spmv_timed(capi.MAT_TANGENT, 20)
print("spmv: %.4f ms/launch, %.3f GB -> %.1f GB/s" % (ms, nbytes / 1e9, nbytes / ms / 1e6))
h.profile(reset=True)
for prec in (capi.PRECOND_BLOCK_JACOBI,):
    h.set_option(capi.OPT_PRECONDITIONER, prec)
    h.nl_begin_step()
    h.nl_newton_assemble()
    t0 = time.time()
    it, res, upd = h.nl_newton_solve(0, 1e-6, 1.0)
    print("precond %d: CG its %d res %.3e upd %.3e wall %.1f ms" % (prec, it, res, upd, (time.time() - t0) * 1e3))
    print(h.profile(reset=True))
part = solvers.FakeParticipant(3, nsteps, p.delta_t, lambda t, it: np.tile([2000.0, 0.0, 0.0], n))
solid = solvers.Solid(prob, part, handle=h)
solid.adapter.initialize(prob)
for s in range(nsteps):
    t0 = time.time()
    solid.step()
    h.synchronize()
    dt = time.time() - t0
    rows = solid.history[-1]
    print("step %d: %.2f s, newton solves %d, cg its %s" % (s, dt, len(rows), [r[0] for r in rows]))
    print("  ", h.profile(reset=True))
