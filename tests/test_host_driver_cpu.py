"""The C++ drop-in host driver (elasticity_2d: Solid / ElastoDynamics / Adapter / Time mirrors of
nonlinear_elasticity.cc:99-167,410-499, linear_elasticity.cc:634-716, adapter.h:346-489) exercised
WITHOUT a GPU: a scripted stand-in for the device library (tests/fake_graftfem.c) is LD_PRELOADed,
logs every C-ABI call and hands out scripted Newton norms. Checks the call sequence of the coupling
loop (explicit and implicit with checkpoints), the Newton control flow against the counts THE
REFERENCE'S OWN solve_nonlinear_timestep produced for the same scripts
(tests/golden/reference_vectors.npz), the error path, and the output files."""
import os
import subprocess

import numpy as np
import pytest

from dealii_adapter_b200 import build

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

FAKE_CFG = """dimensions = 2
time-window-size = 0.01
max-time-windows = {windows}
sub-iterations = {sub}
traction = 0.0,-1500.0
ramp-time = 0.0
watch-point = 0.6,0.2
watch-point-file = watchpoint.log
"""


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    exes = build.build_elasticity()
    d = tmp_path_factory.mktemp("fake_device")
    lib = d / "libfake_graftfem.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(os.path.dirname(HERE), "include"),
                           "-o", str(lib), os.path.join(HERE, "fake_graftfem.c")])
    return exes, str(lib)


def run(env, workdir, prm_text, windows=2, sub=1, res=None, upd=None, extra_env=None):
    exes, lib = env
    (workdir / "parameters.prm").write_text(prm_text)
    (workdir / "precice-config.fake").write_text(FAKE_CFG.format(windows=windows, sub=sub))
    script = workdir / "script.txt"
    script.write_text(" ".join(repr(float(x)) for x in (res or [1.0, 1e-12])) + "\n" +
                      " ".join(repr(float(x)) for x in (upd or [1.0, 1e-12])) + "\n")
    e = dict(os.environ, LD_PRELOAD=lib, GF_FAKE_LOG=str(workdir / "calls.log"),
             GF_FAKE_SCRIPT=str(script), GF_PRECONDITIONER="block-jacobi")
    e.update(extra_env or {})
    r = subprocess.run([exes[0], "parameters.prm"], cwd=workdir, capture_output=True, text=True,
                       timeout=120, env=e)
    calls = [l.split()[0] for l in (workdir / "calls.log").read_text().splitlines()]
    return r, calls


def nl_prm(**repl):
    text = open(os.path.join(GOLDEN, "parameters_nonlinear_fsi3.prm")).read()
    for a, b in repl.items():
        assert a in text, a
        text = text.replace(a, b)
    return text


def test_explicit_coupling_call_sequence_and_output(env, tmp_path):
    # every step: two scripted Newton solves, converged at the third assembly (the update norm of
    # the FIRST solve normalises to 1, so convergence needs a second one, :459-463)
    r, calls = run(env, tmp_path, nl_prm(**{"Output interval       = 10": "Output interval       = 1"}),
                   windows=2, res=[1.0, 1e-3, 1e-12] * 2, upd=[1.0, 1e-9] * 2)
    assert r.returncode == 0, r.stderr
    step = ["nl_begin_step", "set_traction"] + ["nl_newton_assemble", "nl_newton_solve"] * 2 + \
        ["nl_newton_assemble", "nl_end_step", "get_interface_displacement"]
    # no write of initial data: Adapter::initialize only formats it if requiresInitialData()
    # (adapter.h:329-334), which the scripted participant denies
    assert calls == ["create", "postprocess"] + (step + ["postprocess"]) * 2 + ["destroy"]
    assert r.stdout.count("CONVERGED!") == 2
    for k in range(3):      # output_results before the loop and after each completed window
        assert (tmp_path / "dealii-output" / ("solution-%03d.vtk" % k)).exists()
    log = np.loadtxt(tmp_path / "watchpoint.log")
    assert log.shape == (2, 6) and np.allclose(log[:, 4], [1e-3, 2e-3])   # the fake's displacement


def test_implicit_coupling_saves_and_restores_around_every_pass(env, tmp_path):
    r, calls = run(env, tmp_path, nl_prm(), windows=2, sub=3, res=[1.0, 1e-3, 1e-12] * 6,
                   upd=[1.0, 1e-9] * 6)
    assert r.returncode == 0, r.stderr
    body = ["nl_begin_step", "set_traction"] + ["nl_newton_assemble", "nl_newton_solve"] * 2 + \
        ["nl_newton_assemble", "nl_end_step", "get_interface_displacement"]
    window = ["state_save"] + body + ["state_restore"] + body + ["state_restore"] + body
    # save_current_state_if_required first in the loop body (:119), reload after advance (:147)
    assert calls == ["create", "postprocess"] + window * 2 + ["destroy"]
    log = np.loadtxt(tmp_path / "watchpoint.log")     # columns: time, iteration, x, y, ux, uy
    assert log.shape == (6, 6) and log[:, 1].tolist() == [0, 1, 2] * 2
    assert np.allclose(log[:, 0], [0.01] * 3 + [0.02] * 3)     # Time restored with the checkpoint


@pytest.mark.parametrize("k", range(17))
def test_newton_flow_of_the_cxx_mirror_follows_the_reference(env, tmp_path, k):
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    max_it, tol_f, tol_u = ref["newton%02d_in" % k]
    solves, assemblies, converged = (int(x) for x in ref["newton%02d_out" % k])
    prm = nl_prm(**{"Max iterations Newton-Raphson = 10": "Max iterations Newton-Raphson = %d" % int(max_it),
                    "Tolerance force               = 1.0e-9": "Tolerance force               = %r" % float(tol_f),
                    "Tolerance displacement        = 1.0e-6": "Tolerance displacement        = %r" % float(tol_u)})
    r, calls = run(env, tmp_path, prm, windows=1, res=list(ref["newton%02d_res" % k]),
                   upd=list(ref["newton%02d_upd" % k]))
    assert (calls.count("nl_newton_solve"), calls.count("nl_newton_assemble")) == (solves, assemblies)
    if converged:
        assert r.returncode == 0 and "CONVERGED!" in r.stdout
    else:       # AssertThrow(...ExcMessage("No convergence in nonlinear solver!")) -> exit code 1
        assert r.returncode == 1 and "No convergence in nonlinear solver!" in r.stderr
        assert "nl_end_step" not in calls


def test_linear_model_call_sequence(env, tmp_path):
    prm = nl_prm(**{"Model                     = neo-Hookean": "Model                     = linear"})
    r, calls = run(env, tmp_path, prm, windows=3)
    assert r.returncode == 0, r.stderr
    assert calls == ["create", "postprocess", "lin_assemble_once"] + \
        ["set_traction", "lin_step", "get_interface_displacement"] * 3 + ["destroy"]


def expected_calls(events, newton):
    """C-ABI calls the host mirror must issue for the event order of the reference's run()."""
    out, k = [], 0
    while k < len(events):
        e = events[k]
        if e in ("system_setup", "setup_system"):
            out.append("create")
        elif e.startswith("output_results@"):
            out.append("postprocess")
        elif e == "assemble_system":
            out.append("lin_assemble_once")
        elif e == "adapter.save_current_state_if_required:1":
            out.append("state_save")
        elif e == "adapter.reload_old_state_if_required:1":
            out.append("state_restore")
        elif e == "solution_delta=0":
            out.append("nl_begin_step")
        elif e.startswith("adapter.read_data"):
            out.append("set_traction")
        elif e == "solve_nonlinear_timestep":
            out += newton
        elif e == "total_displacement+=solution_delta":
            assert events[k + 1:k + 4] == ["update_acceleration", "update_velocity", "update_old_variables"]
            out.append("nl_end_step")
            k += 3
        elif e == "assemble_rhs":
            assert events[k + 1:k + 3] == ["solve", "update_displacement"]
            out.append("lin_step")
            k += 2
        elif e.startswith("adapter.advance"):
            out.append("get_interface_displacement")
        elif e == "precice.finalize":
            out.append("destroy")
        k += 1
    return out


@pytest.mark.parametrize("k", range(7))
def test_coupling_loop_of_the_cxx_mirror_follows_the_reference_run(env, tmp_path, k):
    """The order of events of the reference's own run() (nonlinear_elasticity.cc:96-167,
    linear_elasticity.cc:632-716; cut out and run against recording stand-ins under a scripted
    coupling scheme, tests/golden/reference_vectors.npz) against the C-ABI call log of the C++ host
    under the same scheme: checkpoint save before the step, restore after advance, output only for
    completed windows at the output interval, the constant-step-size error."""
    ref = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    solver, windows, sub, interval, dt, dt_precice = (str(x) for x in ref["loop%d_case" % k])
    events = [str(e) for e in ref["loop%d_events" % k]]
    prm = nl_prm(**{"Output interval       = 10": "Output interval       = %s" % interval})
    if solver == "lin":
        prm = prm.replace("Model                     = neo-Hookean", "Model                     = linear")
    n_pass = int(windows) * int(sub)
    exes, lib = env
    (tmp_path / "parameters.prm").write_text(prm)
    (tmp_path / "precice-config.fake").write_text(
        FAKE_CFG.format(windows=windows, sub=sub).replace("time-window-size = 0.01",
                                                          "time-window-size = %s" % dt_precice))
    (tmp_path / "script.txt").write_text(" ".join(["1.0", "1e-3", "1e-12"] * n_pass) + "\n" +
                                         " ".join(["1.0", "1e-9"] * n_pass) + "\n")
    e = dict(os.environ, LD_PRELOAD=lib, GF_FAKE_LOG=str(tmp_path / "calls.log"),
             GF_FAKE_SCRIPT=str(tmp_path / "script.txt"), GF_PRECONDITIONER="block-jacobi")
    r = subprocess.run([exes[0], "parameters.prm"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=120, env=e)
    calls = [l.split()[0] for l in (tmp_path / "calls.log").read_text().splitlines()]
    newton = ["nl_newton_assemble", "nl_newton_solve"] * 2 + ["nl_newton_assemble"]
    want = expected_calls(events, newton)
    if int(ref["loop%d_exit" % k]) == 0:
        assert r.returncode == 0, r.stderr
        assert calls == want
        # output file indices = timestep / interval of every output_results event (:1240-1243)
        idx = sorted({int(e.split("@")[1]) // int(interval) for e in events if e.startswith("output_results@")})
        got = sorted(int(f.name[9:12]) for f in (tmp_path / "dealii-output").glob("solution-*.vtk"))
        assert got == idx
    else:
        assert any(e.startswith("THROW:") for e in events)
        assert r.returncode == 1 and "This solver supports only constant time-step sizes" in r.stderr
        assert calls[:len(want)] == want          # everything up to the throw, then the destructor


def test_timer_summary_keeps_the_reference_sections(env, tmp_path):
    """TimerOutput(std::cout, summary, wall_times) (nonlinear_elasticity.cc:79, linear:63): the
    section names of the reference and their call counts, printed when the solver is destroyed."""
    r, calls = run(env, tmp_path, nl_prm(), windows=2, res=[1.0, 1e-3, 1e-12] * 2, upd=[1.0, 1e-9] * 2)
    assert r.returncode == 0, r.stderr
    rows = {l.split("|")[1].strip(): int(l.split("|")[2]) for l in r.stdout.splitlines()
            if l.startswith("| ") and l.count("|") == 5 and l.split("|")[2].strip().isdigit()}
    assert rows == {"Setup system": 1, "Output results": 1, "Assemble linear system": 6,
                    "Linear solver": 4, "Advance adapter": 2}
    assert "Total wallclock time elapsed since start" in r.stdout
    d = tmp_path / "lin"
    d.mkdir()
    prm = nl_prm(**{"Model                     = neo-Hookean": "Model                     = linear"})
    r, calls = run(env, d, prm, windows=3)
    rows = {l.split("|")[1].strip(): int(l.split("|")[2]) for l in r.stdout.splitlines()
            if l.startswith("| ") and l.count("|") == 5 and l.split("|")[2].strip().isdigit()}
    assert rows == {"Output results": 1, "Assemble rhs": 3, "Solve system": 3, "Advance adapter": 3}


def test_shipped_parameter_file_degree_3_reaches_the_device_with_the_right_mesh(env, tmp_path):
    """The reference's own parameters.prm (linear, FSI3, degree 3, Direct): the C++ host builds
    the Q3 mesh (18 x 3 cells -> 55 x 10 nodes), hands it to gf_create and runs the linear call
    sequence; no multigrid attach above degree 2. The participant configuration is read from the
    file name the prm asks for (precice-config.xml)."""
    from test_zz_gpu_high_degree import SHIPPED_PARAMETERS_PRM
    exes, lib = env
    (tmp_path / "parameters.prm").write_text(SHIPPED_PARAMETERS_PRM)
    (tmp_path / "precice-config.xml").write_text(
        FAKE_CFG.format(windows=3, sub=1).replace("time-window-size = 0.01", "time-window-size = 0.005"))
    e = dict(os.environ, LD_PRELOAD=lib, GF_FAKE_LOG=str(tmp_path / "calls.log"))
    r = subprocess.run([exes[0], "parameters.prm"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=120, env=e)
    assert r.returncode == 0, r.stderr
    log = (tmp_path / "calls.log").read_text().splitlines()
    assert log[0].startswith("create dim=2 degree=3 model=0 n_dofs=%d n_cells=54" % (55 * 10 * 2))
    calls = [l.split()[0] for l in log]
    assert "mg_attach" not in calls
    assert calls == ["create", "postprocess", "lin_assemble_once"] + \
        ["set_traction", "lin_step", "get_interface_displacement"] * 3 + ["destroy"]
    assert "Polynomial degree: 3" in r.stdout
