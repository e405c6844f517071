"""SpMV on Q1 rows (81 values per scalar row: BASELINE configs[3] row shape) - launch times of the
kernel kinds on a 96x384x96-cell linear cantilever (10.9 M DoFs, 7.7 GB system matrix), and, with
--ncu, one profiled launch of the default kernel inside a cudaProfilerStart/Stop range."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dealii_adapter_b200 import capi
from dealii_adapter_b200.problem import SolverParameters, make_problem
p = SolverParameters(model="linear", type_lin="CG", poly_degree=1, scenario="PF", delta_t=0.005,
                     mu=0.5e6, nu=0.4, rho=1000.0, theta=0.5, max_iterations_lin=1.0, end_time=1e9)
prob = make_problem(p, 3, reps=[96, 384, 96], numbering="lexicographic")
h = capi.Handle(prob, slab_axis=1)
h.lin_assemble_once()
h.set_vector(capi.VEC_SCRATCH0, np.random.RandomState(1234).uniform(-1, 1, prob.n_dofs))
if "--ncu" in sys.argv:
    rt = ctypes.CDLL("libcudart.so")
    for k in range(2):
        h.spmv(capi.MAT_SYSTEM, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    rt.cudaProfilerStart()
    h.spmv(capi.MAT_SYSTEM, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    rt.cudaProfilerStop()
else:
    for kind in (1, 5, 3, 6, 0):
        h.set_option(capi.OPT_SPMV_KERNEL, kind)
        ms, nbytes = h.spmv_timed(capi.MAT_SYSTEM, 10)
        print(json.dumps({"q1_spmv_kind": kind, "ms": ms, "gbs": nbytes / ms / 1e6, "n_dofs": prob.n_dofs}), flush=True)
h.close()
