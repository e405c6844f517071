"""Sweep of the V-cycle parameters on the bench workload (GPU, 1 rank): Chebyshev degree of the
smoothers, eigenvalue ratio, degree of the coarsest-level solve -> CG iterations and ms per Newton
solve. Usage: python tools/mg_sweep.py [layers]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dealii_adapter_b200 import capi, multigrid, solvers

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 144
prob = bench.make_flap(layers)
H = multigrid.Hierarchy(prob)
h = H.fine
buf = np.tile(bench.TRACTION, h.n_iface_nodes)
fp = solvers.FakeParticipant(3, 10 ** 9, prob.params.delta_t, lambda t, it: buf, bench.N_SUB)
s = solvers.Solid(prob, fp, handle=h)
s.adapter.n_interface_nodes = h.n_iface_nodes
s.adapter.interface_nodes_ids = np.arange(h.n_iface_nodes, dtype=np.int32)
for k in range(8):
    s.step()
h.state_save()                  # every configuration starts from the same state
rows = []
configs = [(3, 40, 80), (3, 40, 60), (3, 40, 40), (3, 40, 30), (3, 20, 60), (3, 30, 60), (3, 60, 60),
           (2, 30, 60), (2, 40, 80), (4, 40, 60), (4, 60, 60), (3, 40, 80)]
if os.environ.get("GF_SWEEP_CONFIGS"):
    configs = [tuple(int(x) for x in c.split(",")) for c in os.environ["GF_SWEEP_CONFIGS"].split(";")]
for deg, ratio, cdeg in configs:
    h.set_option(capi.OPT_MG_SMOOTHER_DEGREE, deg)
    h.set_option(capi.OPT_MG_SMOOTHER_RATIO, ratio)
    h.set_option(capi.OPT_MG_COARSE_DEGREE, cdeg)
    h.state_restore()
    fp.window, fp.iteration = 4, 0
    h0, n0 = len(s.history), s.newton_solves
    h.synchronize(); t0 = time.perf_counter()
    s.step(); s.step()
    h.synchronize(); dt = time.perf_counter() - t0
    solves = s.newton_solves - n0
    its = [r[0] for rows_ in s.history[h0:] for r in rows_]
    row = {"smoother_degree": deg, "ratio": ratio, "coarse_degree": cdeg, "newton_solves": solves,
           "cg_its_total": int(sum(its)), "cg_its": its, "ms_per_solve": 1e3 * dt / solves,
           "dofs_per_s": prob.n_dofs * solves / dt}
    rows.append(row)
    print(json.dumps(row), flush=True)
H.close()
