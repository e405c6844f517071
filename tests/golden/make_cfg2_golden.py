"""Generates tests/golden/cfg2_fsi3_q2.npz: BASELINE.json configs[1] (Turek-Hron FSI3 beam, 2D Q2,
3 global refinements of the 18x3 grid = 144x24 cells, 28,322 DoFs, neo-Hookean, Newmark, dummy
fluid traction ramped over two steps) run through the CPU ORACLE (oracle/oracle.cpp, the
restatement of nonlinear_elasticity.cc:410-499,872-1211) with tight linear solves ("Direct").
The reference itself cannot run here (deal.II/preCICE absent), so this fixture pins the GPU path
to the oracle at the full cfg2 size without re-running the 50 s CPU solve in every GPU test run.

  python tests/golden/make_cfg2_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

N_STEPS = 2
LOAD = (0.0, -200.0)


def cfg2_problem():
    from helpers import nl_params
    from dealii_adapter_b200.problem import make_problem
    p = nl_params(poly_degree=2, scenario="FSI3", type_lin="Direct", delta_t=0.01)
    return make_problem(p, 2, reps=[144, 24])


def traction(prob, step):
    return np.tile(np.array(LOAD) * min(1.0, (step + 1) / 2.0), prob.n_iface_nodes)


def run_oracle(n_steps=N_STEPS):
    from oracle import oracle_py as orc
    prob = cfg2_problem()
    o = orc.Oracle(prob)
    counts, written, res0 = [], [], []
    for s in range(n_steps):
        o.format_precice_to_deal(traction(prob, s), orc.NL_EXTERNAL_STRESS)
        n, hist = o.nl_timestep()
        counts.append(n)
        written.append(o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT))
    return prob, np.array(counts), np.array(written)


if __name__ == "__main__":
    prob, counts, written = run_oracle()
    # watch point of the tutorial: the interface vertex nearest the beam tip (0.6, 0.2)
    pos = np.asarray(prob.interface_positions()).reshape(-1, 2)
    tip = int(np.argmin(np.abs(pos - np.array([0.6, 0.2])).sum(axis=1)))
    np.savez_compressed(os.path.join(HERE, "cfg2_fsi3_q2.npz"), newton_counts=counts,
                        interface_displacement=written, watch_point_index=tip,
                        n_dofs=prob.n_dofs, n_cells=prob.mesh.n_cells)
    print("cfg2 golden: n_dofs", prob.n_dofs, "Newton counts", counts, "tip displacement",
          written[:, 2 * tip:2 * tip + 2])
