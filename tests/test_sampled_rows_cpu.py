"""The sampled-row device of tests/sampled_rows.py, checked on the CPU: on meshes small enough for
the oracle to assemble everything, the rows the SUB-MESH oracle produces at the sampled DoFs must be
the rows of the full-mesh oracle (same cells, same order => bit for bit). This is what lets the GPU
tests compare matrix rows at BASELINE's full sizes (tests/test_gpu_baseline_configs.py)."""
import numpy as np
import pytest

import sampled_rows as sr
from helpers import lin_params, nl_params, smooth_field
from dealii_adapter_b200.problem import make_problem


@pytest.fixture(scope="module")
def orc(native_libs):
    from oracle import oracle_py
    return oracle_py


@pytest.mark.parametrize("dim,degree,reps,numbering,k", [
    (3, 2, [3, 8, 2], "lexicographic", 3), (3, 1, [6, 16, 6], "cellwise", 1),
    (2, 2, [8, 24], "component_wise", 1)])
def test_sub_mesh_rows_equal_full_mesh_rows_nonlinear(orc, dim, degree, reps, numbering, k):
    p = nl_params(poly_degree=degree, body_force=(0.3, -9.81, 0.2 if dim == 3 else 0.0))
    prob = make_problem(p, dim, reps=reps, numbering=numbering)
    u, du = smooth_field(prob, 4e-3, 1), smooth_field(prob, 4e-4, 2)
    v_old, a_old = smooth_field(prob, 0.05, 3), smooth_field(prob, 2.0, 4)
    buf = np.random.RandomState(0).uniform(-2e3, 2e3, prob.n_iface_nodes * dim)
    sample = sr.sample_nodes(prob, n_per_class=k, axis=1)
    assert 10 <= len(sample) < prob.mesh.n_nodes
    got = sr.oracle_nl_rows(orc, prob, sample, u, du, v_old, a_old, buf)
    assert got["n_sub_cells"] < prob.mesh.n_cells
    o = orc.Oracle(prob)
    o.set(orc.NL_TOTAL_DISPLACEMENT, u)
    o.set(orc.NL_SOLUTION_DELTA, du)
    o.set(orc.NL_VELOCITY_OLD, v_old)
    o.set(orc.NL_ACCELERATION_OLD, a_old)
    o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
    assert np.array_equal(o.get(orc.NL_EXTERNAL_STRESS), sr.traction_vector(prob, buf))
    o.nl_update_acceleration()
    o.nl_assemble_system()
    rp, col, val = sr.rows_of(o.csr(orc.MAT_TANGENT), got["rows"], np.arange(prob.n_dofs))
    assert np.array_equal(rp, got["rowptr"]) and np.array_equal(col, got["col"])
    assert np.array_equal(val, got["val"])
    assert np.array_equal(o.get(orc.NL_SYSTEM_RHS)[got["rows"]], got["rhs"])
    # the sample really contains constrained, interface and interior rows
    con = prob.constrained[got["rows"]] != 0
    assert con.any() and (~con).any() and np.isin(got["rows"], prob.iface_dofs.reshape(-1)).any()


@pytest.mark.parametrize("dim,degree,reps", [(3, 1, [6, 16, 6]), (2, 2, [6, 18])])
def test_sub_mesh_rows_equal_full_mesh_rows_linear(orc, dim, degree, reps):
    p = lin_params(poly_degree=degree, type_lin="CG")
    prob = make_problem(p, dim, reps=reps, numbering="lexicographic")
    sample = sr.sample_nodes(prob, n_per_class=1, axis=1)
    got = sr.oracle_lin_rows(orc, prob, sample)
    o = orc.Oracle(prob)
    o.lin_assemble_system()
    o.lin_assemble_rhs()
    for name, which in (("K", orc.MAT_STIFFNESS), ("M", orc.MAT_MASS), ("A", orc.MAT_SYSTEM)):
        rp, col, val = sr.rows_of(o.csr(which), got["rows"], np.arange(prob.n_dofs))
        assert np.array_equal(rp, got["rowptr"]) and np.array_equal(col, got["col"])
        assert np.array_equal(val, got[name]), name


def test_committed_fixtures_are_what_the_oracle_produces(orc):
    """tests/golden/sampled_rows_cfg{3,4}.npz re-derived (cfg3 here; cfg4 needs the 51 M-DoF mesh
    in host memory and is re-derived by the GPU test and by the generator script)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_sampled_rows as g
    gold = np.load(os.path.join(here, "golden", "sampled_rows_cfg3.npz"))
    ref = g.cfg3_oracle_rows()
    for k in ("rows", "rowptr", "col"):
        assert np.array_equal(ref[k], gold[k])
    assert np.array_equal(ref["val"], gold["val"]) and np.array_equal(ref["rhs"], gold["rhs"])
    assert len(gold["rows"]) == 1026
