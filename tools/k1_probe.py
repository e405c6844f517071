"""Direct timing of the cell kernel (K1) + scatter (K3) on the cfg3 mesh: one handle without the
multigrid hierarchy, GF_OPT_PROFILE = 1 around repeated assemblies of a loaded state."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from helpers import smooth_field
from dealii_adapter_b200 import capi
prob = bench.make_flap_reps(bench.CELLS_PER_GPU)
h = capi.Handle(prob)
h.set_vector(capi.NL_TOTAL_DISPLACEMENT, smooth_field(prob, 0.004, 5))
h.set_traction(np.tile(bench.TRACTION, h.n_iface_nodes))
h.nl_begin_step()
for k in range(2):
    h.nl_newton_assemble()
h.set_option(capi.OPT_PROFILE, 1)
h.profile(reset=True)
n = 5
for k in range(n):
    h.nl_newton_assemble()
p = h.profile(reset=True)
cells = prob.mesh.n_cells
ms = p["assemble_cells_ms"] / n
# b <= a node blocks: 378 pairs x 30 FMA x 64 q-points x 2 flops
flops = cells * 378 * 30 * 64 * 2
print(json.dumps({"k1_ms": ms, "scatter_ms": p["scatter_ms"] / n, "cells": cells,
                  "k1_tflops_lower_triangle": flops / ms / 1e9,
                  "k1_frac_of_dfma_peak_33.9": flops / ms / 1e9 / 33.9}))
h.close()
