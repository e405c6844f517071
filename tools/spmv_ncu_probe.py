"""Minimal target for an ncu capture of the SpMV variants on the cfg3 tangent: one assembly, then
inside a cudaProfilerStart/Stop range one launch each of
  kernel 0 (single ring) FP64, kernel 0 FP32 copy, kernel 2 (two rings) FP64, kernel 2 FP32 copy.
Run:  ncu --set full --profile-from-start off --import-source on -o out python tools/spmv_ncu_probe.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dealii_adapter_b200 import capi, multigrid  # noqa: E402

prob = bench.make_flap_reps(bench.CELLS_PER_GPU)
H = multigrid.Hierarchy(prob, n_levels=2)
h = H.fine
h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
h.set_traction(np.tile(bench.TRACTION, h.n_iface_nodes))
h.nl_begin_step()
h.nl_newton_assemble()
rng = np.random.RandomState(1234)
h.set_vector(capi.VEC_SCRATCH0, rng.uniform(-1, 1, prob.n_dofs))
rt = ctypes.CDLL("libcudart.so")
for kind in (0, 2):          # warm-up outside the range
    h.set_option(capi.OPT_SPMV_KERNEL, kind)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
rt.cudaProfilerStart()
for kind in (0, 2):
    h.set_option(capi.OPT_SPMV_KERNEL, kind)
    h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
    h.spmv(capi.MAT_MG_F32, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
rt.cudaProfilerStop()
H.close()
print("done")
