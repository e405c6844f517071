"""SpMV tuning on the bench matrix (GPU, 1 rank): L2 prefetch distance x gather mode of the
TMA-tiled kernel -> ms per launch, GB/s of the stored format; results must stay bitwise equal."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dealii_adapter_b200 import capi

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 144
prob = bench.make_flap(layers)
h = capi.Handle(prob)
h.set_traction(np.tile(bench.TRACTION, h.n_iface_nodes))
h.nl_begin_step()
h.nl_newton_assemble()
x = np.random.RandomState(1234).uniform(-1, 1, prob.n_dofs)
h.set_vector(capi.VEC_SCRATCH0, x)
ref = None
for gather in (0, 1):
    for pf in (0, 1, 2, 4, 8, 16, 32):
        h.set_option(capi.OPT_SPMV_GATHER, gather)
        h.set_option(capi.OPT_SPMV_PREFETCH, pf)
        h.spmv(capi.MAT_TANGENT, capi.VEC_SCRATCH0, capi.VEC_SCRATCH1)
        y = h.get_vector(capi.VEC_SCRATCH1)
        if ref is None:
            ref = y
        ms, nbytes = h.spmv_timed(capi.MAT_TANGENT, 30)
        print(json.dumps({"gather": gather, "prefetch_tiles": pf, "ms": ms, "GBps": nbytes / ms / 1e6,
                          "bitwise_equal": bool(np.array_equal(y, ref))}), flush=True)
h.set_option(capi.OPT_SPMV_KERNEL, 1)
ms, nbytes = h.spmv_timed(capi.MAT_TANGENT, 30)
print(json.dumps({"kernel": "ldg", "ms": ms, "GBps": nbytes / ms / 1e6}))
h.close()
