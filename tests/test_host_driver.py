"""The C++ drop-in host driver (elasticity_2d/3d: Solid / ElastoDynamics / Adapter / Parameters /
Time mirrors + scripted precice::Participant). CPU tests cover the parameter parser and error
behaviour (exit code 1 + message, elasticity.cc:101-126); the gpu test runs a coupled simulation and
compares the watch-point log with the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import nl_params, rel_err
from dealii_adapter_b200 import build
from dealii_adapter_b200.problem import make_problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FAKE_CFG = """dimensions = 2
time-window-size = 0.01
max-time-windows = 4
sub-iterations = 1
traction = 0.0,-1500.0
ramp-time = 0.03
watch-point = 0.6,0.2
watch-point-file = watchpoint.log
"""


@pytest.fixture(scope="module")
def exes():
    return build.build_elasticity()


def run(exe, workdir, prm="parameters.prm"):
    return subprocess.run([exe, prm], cwd=workdir, capture_output=True, text=True, timeout=600)


def test_unknown_key_and_bad_value_abort_with_exit_code_1(exes, tmp_path):
    text = open(os.path.join(GOLDEN, "parameters_nonlinear_fsi3.prm")).read()
    (tmp_path / "parameters.prm").write_text(text.replace("set rho ", "set density "))
    r = run(exes[0], tmp_path)
    assert r.returncode == 1 and "Exception on processing" in r.stderr and "density" in r.stderr
    (tmp_path / "parameters.prm").write_text(text.replace("= 0.4", "= 0.7"))   # nu in [-1,0.5]
    r = run(exes[0], tmp_path)
    assert r.returncode == 1 and "Poisson" in r.stderr
    (tmp_path / "parameters.prm").write_text(text.replace("= Stress", "= Pressure"))
    r = run(exes[0], tmp_path)
    assert r.returncode == 1 and "Unknown read data type" in r.stderr
    # the stale subsection names of nonlinear_elasticity.prm are rejected as well
    (tmp_path / "parameters.prm").write_text(text.replace("subsection Solver", "subsection Linear solver"))
    r = run(exes[0], tmp_path)
    assert r.returncode == 1 and "no such subsection" in r.stderr


def test_force_data_rejected_by_neo_hookean_and_dimension_mismatch(exes, tmp_path):
    text = open(os.path.join(GOLDEN, "parameters_nonlinear_fsi3.prm")).read()
    (tmp_path / "precice-config.fake").write_text(FAKE_CFG)
    (tmp_path / "parameters.prm").write_text(text.replace("= Stress", "= Force"))
    r = run(exes[0], tmp_path)
    assert r.returncode == 1 and "doesn't support 'Force' data" in r.stderr
    (tmp_path / "parameters.prm").write_text(text)
    (tmp_path / "precice-config.fake").write_text(FAKE_CFG.replace("dimensions = 2", "dimensions = 3"))
    r = run(exes[0], tmp_path)
    # without a GPU the device set-up fails first; with one, the dimension check of
    # Adapter::initialize (adapter.h:235-240) fires. Either way: exit code 1, message on stderr.
    assert r.returncode == 1 and "Exception on processing" in r.stderr


@pytest.mark.gpu
def test_coupled_nonlinear_run_matches_oracle_watchpoint(exes, tmp_path, native_libs):
    from oracle import oracle_py as orc
    shutil.copy(os.path.join(GOLDEN, "parameters_nonlinear_fsi3.prm"), tmp_path / "parameters.prm")
    (tmp_path / "precice-config.fake").write_text(FAKE_CFG)
    r = run(exes[0], tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    assert "CONVERGED!" in r.stdout and "LIN_IT" in r.stdout
    # output_results before the first step (:103): DataOut patches on the (still undisplaced) grid,
    # one Lagrange quadrilateral of order 2 per cell, 18 x 3 cells
    assert "Output written to solution-000.vtk" in r.stdout
    vtk = (tmp_path / "dealii-output" / "solution-000.vtk").read_text()
    assert "POINTS %d double" % (54 * 9) in vtk and "SCALARS strain_xy double 1" in vtk
    log = np.loadtxt(tmp_path / "watchpoint.log")
    assert log.shape == (4, 6)
    # oracle: same scenario (FSI3 2D Q2 18x3 cells), Direct stand-in, ramped traction
    p = nl_params(poly_degree=2, scenario="FSI3", type_lin="Direct", delta_t=0.01,
                  max_iterations_lin=2.0)
    prob = make_problem(p, 2, numbering="component_wise")
    o = orc.Oracle(prob)
    pos = prob.interface_positions().reshape(-1, 2)
    k = np.argmin(((pos - np.array([0.6, 0.2])) ** 2).sum(axis=1))
    assert np.allclose(log[0, 2:4], pos[k])
    counts = []
    for step in range(4):
        t = (step + 1) * 0.01
        o.format_precice_to_deal(np.tile(np.array([0.0, -1500.0]) * min(1.0, t / 0.03),
                                         prob.n_iface_nodes), orc.NL_EXTERNAL_STRESS)
        n, _ = o.nl_timestep()
        counts.append(n)
        d = o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT).reshape(-1, 2)[k]
        assert rel_err(log[step, 4:6], d) < 1e-8
    # Newton iteration counts: one table row per linear solve
    rows = [l for l in r.stdout.splitlines() if " SLV " in l]
    assert len(rows) == sum(counts)


@pytest.mark.gpu
def test_coupled_run_on_refined_mesh_uses_multigrid_and_matches_oracle(exes, tmp_path, native_libs):
    """FSI3 2D with one global refinement (36 x 6 cells): the C++ host links the two levels with
    gf_mg_attach and the CG is multigrid-preconditioned; watch-point vs the oracle (tight solves)."""
    from oracle import oracle_py as orc
    shutil.copy(os.path.join(GOLDEN, "parameters_nonlinear_fsi3.prm"), tmp_path / "parameters.prm")
    (tmp_path / "precice-config.fake").write_text(FAKE_CFG + "mesh-repetitions = 36,6\n")
    r = run(exes[0], tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    assert "geometric multigrid, 2 levels" in r.stdout
    log = np.loadtxt(tmp_path / "watchpoint.log")
    p = nl_params(poly_degree=2, scenario="FSI3", type_lin="Direct", delta_t=0.01,
                  max_iterations_lin=2.0)
    prob = make_problem(p, 2, reps=[36, 6], numbering="component_wise")
    o = orc.Oracle(prob)
    pos = prob.interface_positions().reshape(-1, 2)
    k = np.argmin(((pos - np.array([0.6, 0.2])) ** 2).sum(axis=1))
    counts = []
    for step in range(4):
        t = (step + 1) * 0.01
        o.format_precice_to_deal(np.tile(np.array([0.0, -1500.0]) * min(1.0, t / 0.03),
                                         prob.n_iface_nodes), orc.NL_EXTERNAL_STRESS)
        n, _ = o.nl_timestep()
        counts.append(n)
        d = o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT).reshape(-1, 2)[k]
        assert rel_err(log[step, 4:6], d) < 1e-8
    rows = [l for l in r.stdout.splitlines() if " SLV " in l]
    assert len(rows) == sum(counts)
    # block-Jacobi on request
    env = dict(os.environ, GF_PRECONDITIONER="block-jacobi")
    r2 = subprocess.run([exes[0], "parameters.prm"], cwd=tmp_path, capture_output=True, text=True,
                        timeout=600, env=env)
    assert r2.returncode == 0 and "CG preconditioner: block-Jacobi" in r2.stdout
    assert rel_err(np.loadtxt(tmp_path / "watchpoint.log")[:, 4:6], log[:, 4:6]) < 1e-8
