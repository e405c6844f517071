"""Target of the round-2 `ncu --set full` capture (profile range via cudaProfilerStart/Stop, run
with --profile-from-start off): on the cfg3 tangent (24x144x24 Q2 cells, loaded state)
  1. one launch of the default SpMV (GF_OPT_SPMV_KERNEL = 0 -> two-ring kernel, 16 consumer warps,
     transposed row reduction) - the roofline kernel of bench.py, `traffic` comes from here;
  2. one launch of the all-FP32 V-cycle operator (GF_OPT_MG_MATRIX_PRECISION = 2);
  3. one assembly: nl_cells_kernel<3,2> (K1, lower node blocks), scatter_matrix_kernel (K3), ...
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import smooth_field  # noqa: E402
from dealii_adapter_b200 import capi, multigrid  # noqa: E402

prob = bench.make_flap_reps(bench.CELLS_PER_GPU)
H = multigrid.Hierarchy(prob, n_levels=2)
h = H.fine
h.set_option(capi.OPT_MG_MATRIX_PRECISION, 2)
h.set_vector(capi.NL_TOTAL_DISPLACEMENT, smooth_field(prob, 0.004, 5))
h.set_traction(np.tile(bench.TRACTION, h.n_iface_nodes))
h.nl_begin_step()
h.nl_newton_assemble()
h.nl_newton_assemble()
h.set_vector(capi.VEC_SCRATCH0, np.random.RandomState(1234).uniform(-1, 1, prob.n_dofs))
for k in range(2):
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
H.handles[1].set_option(capi.OPT_PROFILE, 0)
h.nl_newton_assemble()
rt.cudaProfilerStop()
H.close()
print("done")
