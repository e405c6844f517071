"""Development check of the multigrid-preconditioned CG against block-Jacobi CG (GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dealii_adapter_b200 import capi, solvers, multigrid as mg
from dealii_adapter_b200.problem import SolverParameters, make_problem

reps = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "12,72,12").split(",")]
model = sys.argv[2] if len(sys.argv) > 2 else "neo-Hookean"
degree = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = SolverParameters(model=model, type_lin="CG", poly_degree=degree, scenario="PF", delta_t=0.01,
                     mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6, max_iterations_lin=1.0, end_time=1e9)
prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
n = prob.n_iface_nodes
buf = np.tile([2000.0, 0.0, 0.0], n)
cls = solvers.Solid if model == "neo-Hookean" else solvers.ElastoDynamics
res = {}
for kind in (("mg",) if os.environ.get("MG_ONLY") else ("jacobi", "mg")):
    t0 = time.time()
    if kind == "mg":
        H = mg.Hierarchy(prob)
        h = H.fine
        print("levels:", [q.mesh.reps for q in H.problems])
    else:
        h = capi.Handle(prob)
    fp = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: buf)
    s = cls(prob, fp, handle=h)
    s.adapter.initialize(prob)
    if model != "neo-Hookean":
        h.lin_assemble_once()
    h.set_option(capi.OPT_PROFILE, 1)
    h.synchronize(); t1 = time.time()
    for k in range(2):
        s.step()
    h.synchronize(); t2 = time.time()
    res[kind] = fp.written[-1][2]
    hist = [[r[0] for r in rows] for rows in s.history] if model == "neo-Hookean" else s.history
    print(kind, "setup %.2fs run %.2fs" % (t1 - t0, t2 - t1), "cg its:", hist)
    print("   profile", {k: round(v, 1) if isinstance(v, float) else v for k, v in h.profile().items() if v})
if "jacobi" not in res: sys.exit(0)
err = np.abs(res["mg"] - res["jacobi"]).max() / np.abs(res["jacobi"]).max()
print("rel diff of interface displacement mg vs jacobi: %.3e" % err)
