"""CPU: the committed cfg2 golden fixture (tests/golden/cfg2_fsi3_q2.npz, made by
tests/golden/make_cfg2_golden.py) is what the oracle produces — first timestep re-run here."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_reproduces_cfg2_golden_first_step(native_libs):
    import make_cfg2_golden as g
    gold = np.load(os.path.join(HERE, "golden", "cfg2_fsi3_q2.npz"))
    prob, counts, written = g.run_oracle(n_steps=1)
    assert prob.n_dofs == int(gold["n_dofs"]) == 28322 and prob.mesh.n_cells == 3456
    assert counts[0] == gold["newton_counts"][0]
    ref = gold["interface_displacement"][0]
    assert np.abs(written[0] - ref).max() <= 1e-12 * np.abs(ref).max()
