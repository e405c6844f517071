"""Generates tests/golden/sampled_rows_cfg3.npz and sampled_rows_cfg4.npz: ORACLE matrix rows and
right-hand-side entries at ~1000 sampled DoFs of BASELINE.json configs[2] (PF 3D Q2 neo-Hookean,
24x144x24 cells, 2,081,667 DoFs) and configs[3] (3D cantilever Q1 linear, 128x1024x128 cells,
51,171,075 DoFs) - the sizes bench.py runs. The oracle assembles the sub-mesh of all cells touching
the sampled nodes (tests/sampled_rows.py; tests/test_sampled_rows_cpu.py proves sub-mesh rows ==
full-mesh rows bit for bit), so the rows are what it would produce on the full mesh. Sampled classes:
interior, faces / edges / corners of the box (clamped, z-clamped, interface), and the node planes
where slab partitions into 2, 4, 8 ranks cut.

The GPU tests (tests/test_gpu_baseline_configs.py, tests/test_gpu_baseline_cfg4.py) re-run these
functions live (the oracle library travels to the GPU box; seconds) AND compare with the committed
files.

  python tests/golden/make_sampled_rows.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CFG3_REPS = [24, 144, 24]
CFG4_REPS = [128, 1024, 128]
CFG3_TRACTION = (1500.0, 0.0, 50.0)
CFG4_TRACTION = (200.0, -30.0, 10.0)


def cfg3_problem():
    from helpers import nl_params
    from dealii_adapter_b200.problem import make_problem
    p = nl_params(poly_degree=2, scenario="PF", type_lin="CG", delta_t=0.01, max_iterations_lin=1.0)
    return make_problem(p, 3, reps=CFG3_REPS, numbering="lexicographic")


def cfg3_state(prob):
    """A loaded, moving state: what a Newton pass in the middle of a run sees."""
    from helpers import smooth_field
    return {"u": smooth_field(prob, 0.004, 5), "du": smooth_field(prob, 0.0004, 8),
            "v_old": smooth_field(prob, 0.05, 6), "a_old": smooth_field(prob, 2.0, 7),
            "traction": np.tile(CFG3_TRACTION, prob.n_iface_nodes)}


def cfg3_oracle_rows(prob=None, n_per_class=8):
    import sampled_rows as sr
    from oracle import oracle_py as orc
    prob = prob or cfg3_problem()
    s = cfg3_state(prob)
    sample = sr.sample_nodes(prob, n_per_class=n_per_class, axis=1)
    return sr.oracle_nl_rows(orc, prob, sample, s["u"], s["du"], s["v_old"], s["a_old"], s["traction"])


def cfg4_problem(reps=None):
    from helpers import lin_params
    from dealii_adapter_b200.problem import make_problem
    p = lin_params(poly_degree=1, type_lin="CG", delta_t=0.005, max_iterations_lin=1.0)
    return make_problem(p, 3, reps=reps or CFG4_REPS, numbering="lexicographic")


def cfg4_oracle_rows(prob=None, n_per_class=8):
    import sampled_rows as sr
    from oracle import oracle_py as orc
    prob = prob or cfg4_problem()
    sample = sr.sample_nodes(prob, n_per_class=n_per_class, axis=1)
    return sr.oracle_lin_rows(orc, prob, sample)


if __name__ == "__main__":
    r3 = cfg3_oracle_rows()
    np.savez_compressed(os.path.join(HERE, "sampled_rows_cfg3.npz"), **r3)
    print("cfg3: %d rows, %d entries, %d sub-mesh cells" % (len(r3["rows"]), len(r3["val"]), r3["n_sub_cells"]))
    r4 = cfg4_oracle_rows()
    np.savez_compressed(os.path.join(HERE, "sampled_rows_cfg4.npz"), **r4)
    print("cfg4: %d rows, %d entries, %d sub-mesh cells" % (len(r4["rows"]), len(r4["K"]), r4["n_sub_cells"]))
