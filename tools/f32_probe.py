"""One-GPU probe of GF_OPT_MG_MATRIX_PRECISION on the bench workload (cfg3): SpMV launch time of
the FP64 matrix vs its FP32 copy, and coupled-step throughput with the V-cycle streaming either.
Prints one JSON object. No torch import (single GPU, ctypes only)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dealii_adapter_b200 import capi, multigrid, solvers  # noqa: E402


def main():
    reps = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else bench.CELLS_PER_GPU
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    t0 = time.perf_counter()
    prob = bench.make_flap_reps(reps)
    H = multigrid.Hierarchy(prob)
    h = H.fine
    buf = np.tile(bench.TRACTION, h.n_iface_nodes)
    part = solvers.FakeParticipant(3, 10 ** 9, prob.params.delta_t, lambda t, it: buf, bench.N_SUB)
    solid = solvers.Solid(prob, part, handle=h)
    solid.adapter.n_interface_nodes = h.n_iface_nodes
    solid.adapter.interface_nodes_ids = np.arange(h.n_iface_nodes, dtype=np.int32)
    out = {"reps": list(reps), "n_dofs": prob.n_dofs, "setup_s": time.perf_counter() - t0}

    def run(prec):
        h.set_option(capi.OPT_MG_MATRIX_PRECISION, prec)
        for k in range(bench.N_SUB):
            solid.step()
        s0, h0 = solid.newton_solves, len(solid.history)
        h.synchronize()
        h.event_record(0)
        for k in range(n_steps):
            solid.step()
        h.event_record(1)
        h.synchronize()
        ms = h.event_elapsed_ms(0, 1)
        solves = solid.newton_solves - s0
        its = [int(r[0]) for rows in solid.history[h0:] for r in rows]
        return {"dofs_per_s": prob.n_dofs * solves / (1e-3 * ms), "ms_per_step": ms / n_steps,
                "newton_solves": solves, "cg_iterations": its,
                "iface_disp_max": float(np.abs(part.written[-1][2]).max())}

    kinds = [int(k) for k in os.environ.get("GF_PROBE_KINDS", "0,2").split(",")]
    for kind in kinds:
        h.set_option(capi.OPT_SPMV_KERNEL, kind)
        r64 = run(0)
        r32 = run(1)
        ms64, b64 = h.spmv_timed(capi.MAT_TANGENT, 20)
        ms32, b32 = h.spmv_timed(capi.MAT_MG_F32, 20)
        out["spmv_kernel_%d" % kind] = {
            "fp64_vcycle": r64, "fp32_vcycle": r32,
            "spmv_fp64": {"ms": ms64, "bytes": b64, "gbs": b64 / ms64 / 1e6},
            "spmv_fp32_copy": {"ms": ms32, "bytes": b32, "gbs": b32 / ms32 / 1e6}}
    print(json.dumps(out))
    H.close()


if __name__ == "__main__":
    main()
