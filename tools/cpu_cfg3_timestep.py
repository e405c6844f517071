#!/usr/bin/env python
"""One-off, offline: ONE coupled timestep of BASELINE configs[2] at its FULL size (24x144x24 Q2
cells, 2,081,667 DoFs) on the CPU through the oracle (the restatement of the reference's path:
threaded cell assembly, serial SolverCG + SSOR(0.65) as deal.II runs it), i.e. the SAME config the
GPU bench line is measured on. Hours of CPU time: not part of bench.py or the tests. Every Newton
solve is logged as it finishes (profiles/r02_cpu_cfg3_full_timestep.jsonl), so a cut-off run still
leaves a same-config anchor for the headline ratio.

  nohup python tools/cpu_cfg3_timestep.py &
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "r02_cpu_cfg3_full_timestep.jsonl")
reps = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["24", "144", "24"])]
prob = bench.make_flap_reps(reps, numbering="cellwise")
p = prob.params
t0 = time.time()
o = orc.Oracle(prob)
free = prob.constrained == 0


def log(**kw):
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")


log(event="start", reps=reps, n_dofs=prob.n_dofs, threads=o.n_threads, create_s=time.time() - t0,
    cpu=os.popen("grep -m1 'model name' /proc/cpuinfo").read().strip())
o.format_precice_to_deal(np.tile(bench.TRACTION, prob.n_iface_nodes), orc.NL_EXTERNAL_STRESS)
o.set(orc.NL_SOLUTION_DELTA, np.zeros(prob.n_dofs))
res0 = u0 = None
t_step = time.time()
solves = 0
for it in range(p.max_iterations_NR):            # nonlinear_elasticity.cc:436-495
    ta = time.time()
    o.nl_update_acceleration()
    o.nl_assemble_system()
    res = o.nl_error_residual()
    t_asm = time.time() - ta
    res0 = res if it == 0 else res0
    rn = res / res0 if res0 != 0 else res
    if it > 0 and ((un <= p.tol_u or u <= 1e-15) and (rn <= p.tol_f or res <= 5e-9)):
        log(event="converged", newton_iterations=it, res_norm=rn, assemble_s=t_asm)
        break
    ts = time.time()
    st, lin_it, lin_res = o.nl_solve_linear_system()
    t_sol = time.time() - ts
    du = o.get(orc.NL_NEWTON_UPDATE)
    u = float(np.linalg.norm(du[free]))
    u0 = u if it == 0 else u0
    un = u / u0 if u0 != 0 else u
    o.set(orc.NL_SOLUTION_DELTA, o.get(orc.NL_SOLUTION_DELTA) + du)
    solves += 1
    log(event="newton_solve", it=it, assemble_s=t_asm, cg_s=t_sol, cg_iterations=lin_it, cg_status=st,
        res_norm=rn, u_norm=un, elapsed_s=time.time() - t_step,
        dofs_per_s_so_far=prob.n_dofs * solves / (time.time() - t_step))
log(event="done", newton_solves=solves, seconds=time.time() - t_step,
    newton_step_dofs_per_s=prob.n_dofs * solves / (time.time() - t_step))
