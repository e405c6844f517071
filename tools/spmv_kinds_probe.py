"""Seconds-long probe of the SpMV kernel kinds (GF_OPT_SPMV_KERNEL 0, 1, 3, 5, 6): bitwise equality of
y = A x against kind 0 on a small 3D Q2 and a 2D Q2 problem, then launch times on the cfg3 tangent
(FP64 and FP32 copy). Prints one JSON line per stage so that a cut-off run still reports."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import nl_params, smooth_field  # noqa: E402
from dealii_adapter_b200 import capi, multigrid  # noqa: E402
from dealii_adapter_b200.problem import make_problem  # noqa: E402

KINDS = [int(k) for k in os.environ.get("GF_PROBE_KINDS", "0,1,3,5,6").split(",")]


def assembled(prob, n_levels=2):
    H = multigrid.Hierarchy(prob, n_levels=n_levels)
    h = H.fine
    h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
    h.set_vector(capi.NL_TOTAL_DISPLACEMENT, smooth_field(prob, 2e-3, seed=2))
    h.nl_begin_step()
    h.nl_newton_assemble()
    h.set_vector(capi.VEC_SCRATCH0, np.random.RandomState(3).uniform(-1, 1, prob.n_dofs))
    return H, h


for dim, reps in ((3, [6, 12, 4]), (2, [16, 24])):
    prob = make_problem(nl_params(poly_degree=2, type_lin="CG"), dim, reps=reps)
    H, h = assembled(prob)
    res = {}
    for mat, tag in ((capi.MAT_TANGENT, "fp64"), (capi.MAT_MG_F32, "fp32")):
        ref = None
        for kind in KINDS:
            h.set_option(capi.OPT_SPMV_KERNEL, kind)
            h.spmv(mat, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
            y = h.get_vector(capi.VEC_SCRATCH1)
            ref = y if ref is None else ref
            res["%s_kind%d_bitwise_equal" % (tag, kind)] = bool(np.array_equal(y, ref))
    print(json.dumps({"stage": "equality", "dim": dim, "reps": reps, **res}), flush=True)
    H.close()

prob = bench.make_flap_reps(bench.CELLS_PER_GPU)
H = multigrid.Hierarchy(prob, n_levels=2)
h = H.fine
h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
h.set_traction(np.tile(bench.TRACTION, h.n_iface_nodes))
h.nl_begin_step()
h.nl_newton_assemble()
for kind in KINDS:
    h.set_option(capi.OPT_SPMV_KERNEL, kind)
    ms64, b64 = h.spmv_timed(capi.MAT_TANGENT, 10)
    ms32, b32 = h.spmv_timed(capi.MAT_MG_F32, 10)
    print(json.dumps({"stage": "cfg3_timing", "kind": kind, "fp64_ms": ms64, "fp64_gbs": b64 / ms64 / 1e6,
                      "fp32_ms": ms32, "fp32_gbs": b32 / ms32 / 1e6}), flush=True)
# all-FP32 operator (GF_OPT_MG_MATRIX_PRECISION = 2), 8+16 warps
h.set_option(capi.OPT_SPMV_KERNEL, 0)
h.set_option(capi.OPT_MG_MATRIX_PRECISION, 2)
h.nl_newton_assemble()
ms32x, b32 = h.spmv_timed(capi.MAT_MG_F32, 10)
print(json.dumps({"stage": "cfg3_timing_all_fp32", "ms": ms32x, "gbs": b32 / ms32x / 1e6}), flush=True)
H.close()
