"""Development timing of the matrix-free operator against the assembled SpMV (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dealii_adapter_b200 import capi
from dealii_adapter_b200.problem import SolverParameters, make_problem
reps = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "24,144,24").split(",")]
p = SolverParameters(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF", delta_t=0.01,
                     mu=0.5e6, nu=0.4, rho=1000.0)
prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
h = capi.Handle(prob)
for op in (0, 1):
    h.set_option(capi.OPT_OPERATOR, op)
    h.nl_begin_step()
    h.set_option(capi.OPT_PROFILE, 1)
    h.profile(reset=True)
    h.nl_newton_assemble()
    pr = h.profile(reset=True)
    h.set_option(capi.OPT_PROFILE, 0)
    ms, nbytes = h.spmv_timed(capi.MAT_TANGENT, 20)
    print("operator %d: assemble cells %.2f ms scatter %.2f ms | apply %.4f ms, %.3f GB -> %.0f GB/s"
          % (op, pr["assemble_cells_ms"], pr["scatter_ms"], ms, nbytes / 1e9, nbytes / ms / 1e6))
