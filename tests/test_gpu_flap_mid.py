"""Coupled run of the bench workload on a mid-size flap (12x48x12 Q2 cells, 181,875 DoFs) with the
bench's EXACT options - multigrid-preconditioned device CG at relative tolerance 1e-6, implicit
coupling k = 2 with checkpoint/restore - against the ORACLE run with the reference's own solver
(SolverCG + SSOR(0.65) at 1e-6, nonlinear_elasticity.cc:1174-1187; committed fixture
tests/golden/flap_mid_q2.npz, generator make_flap_mid_golden.py, ~20 min of CPU):
identical Newton iteration counts, interface and tip watch-point displacement to 1e-8."""
import os
import sys

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.mark.parametrize("mg_precision", [0, 2])
def test_flap_mid_newton_counts_and_watchpoint_match_oracle(native_libs, mg_precision):
    from dealii_adapter_b200 import capi, multigrid, solvers
    native_libs.build_cuda()
    import make_flap_mid_golden as g
    gold = np.load(os.path.join(HERE, "golden", "flap_mid_q2.npz"))
    prob = g.flap_problem()
    assert prob.n_dofs == int(gold["n_dofs"]) == 181875
    H = multigrid.Hierarchy(prob)
    assert H.n_levels == 3
    H.fine.set_option(capi.OPT_MG_MATRIX_PRECISION, mg_precision)
    part = solvers.FakeParticipant(3, g.N_WINDOWS, prob.params.delta_t,
                                   lambda t, it: g.traction(prob, int(round(t / 0.01)) - 1, it), g.N_SUB)
    solid = solvers.Solid(prob, part, handle=H.fine)
    solid.run()
    assert [len(r) for r in solid.history] == list(gold["newton_counts"])
    tip = int(gold["watch_point_index"])
    assert len(part.written) == len(gold["interface_displacement"]) == g.N_WINDOWS * g.N_SUB
    worst = 0.0
    for (w, it, data), ref in zip(part.written, gold["interface_displacement"]):
        worst = max(worst, rel_err(data, ref))
        assert rel_err(data, ref) < 1e-8
        assert abs(data[3 * tip] - ref[3 * tip]) <= 1e-8 * abs(ref[3 * tip])
    cg = [[int(r[0]) for r in rows] for rows in solid.history]
    print("flap_mid (V-cycle precision %d): Newton counts %s, device CG its %s (oracle SSOR-CG: %s), "
          "max interface error %.2e" % (mg_precision, [len(r) for r in solid.history], cg,
                                        gold["oracle_cg_iterations"][:, :4].tolist(), worst))
    H.close()
