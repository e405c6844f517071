"""Generates tests/golden/flap_mid_q2.npz: the bench workload (bench.py: PF 3D Q2 neo-Hookean,
Newmark, implicit coupling k=2 with checkpoint/restore, CG relative tolerance 1e-6 "Residual",
max_iterations_lin 1.0, tol_f 1e-9, tol_u 1e-6) on a mid-size flap - 12x48x12 cells, 181,875 DoFs -
run through the CPU ORACLE with ITS solver (SolverCG + SSOR(0.65), nonlinear_elasticity.cc:1174-1187).
The GPU test (tests/test_gpu_flap_mid.py) runs the same coupled windows with the bench's exact
options (multigrid-preconditioned device CG at 1e-6) and asserts identical Newton counts and
interface / watch-point displacement to 1e-8. The reference itself cannot run here (deal.II/preCICE
absent); the oracle takes ~20 min of CPU for this, hence a committed fixture.

  python tests/golden/make_flap_mid_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

REPS = [12, 48, 12]
N_WINDOWS, N_SUB = 2, 2
LOAD = (2000.0, 0.0, 0.0)


def flap_problem():
    from dealii_adapter_b200.problem import SolverParameters, make_problem
    p = SolverParameters(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF",
                         delta_t=0.01, mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6,
                         max_iterations_lin=1.0, max_iterations_NR=10, tol_f=1e-9, tol_u=1e-6,
                         end_time=1e9)
    return make_problem(p, 3, reps=REPS, numbering="lexicographic")


def traction(prob, window, iteration):
    """Dummy fluid: ramped over the windows, and the first sub-iteration of a window sees a
    different (under-relaxed) load than the accepted one, so the checkpoint restore matters."""
    scale = min(1.0, (window + 1) / 2.0) * (0.8 if iteration == 0 else 1.0)
    return np.tile(np.array(LOAD) * scale, prob.n_iface_nodes)


def run_oracle():
    from oracle import oracle_py as orc
    prob = flap_problem()
    o = orc.Oracle(prob)
    counts, cg_its, written = [], [], []
    for w in range(N_WINDOWS):
        for it in range(N_SUB):
            t0 = time.time()
            o.format_precice_to_deal(traction(prob, w, it), orc.NL_EXTERNAL_STRESS)
            if it == 0:
                o.save_state()
            n, hist = o.nl_timestep()
            counts.append(n)
            cg_its.append([int(r[0]) for r in hist])
            written.append(o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT))
            if it != N_SUB - 1:
                o.reload_state()
            print("window %d it %d: %d Newton solves, CG its %s, %.0f s" % (w, it, n, cg_its[-1],
                                                                          time.time() - t0), flush=True)
    return prob, np.array(counts), cg_its, np.array(written)


if __name__ == "__main__":
    prob, counts, cg_its, written = run_oracle()
    pos = np.asarray(prob.interface_positions()).reshape(-1, 3)
    tip = int(np.argmin(np.abs(pos - np.array([0.0, 1.0, 0.15])).sum(axis=1)))
    np.savez_compressed(os.path.join(HERE, "flap_mid_q2.npz"), newton_counts=counts,
                        oracle_cg_iterations=np.array([x + [0] * (12 - len(x)) for x in cg_its]),
                        interface_displacement=written, watch_point_index=tip,
                        n_dofs=prob.n_dofs, n_cells=prob.mesh.n_cells)
    print("flap_mid golden: n_dofs", prob.n_dofs, "Newton counts", counts, "tip",
          written[:, 3 * tip:3 * tip + 3])
