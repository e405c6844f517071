"""bench.py's control flow on a machine without a GPU: the device handle is replaced by a scripted
stand-in (no arithmetic of the hot path happens here), the real Solid / ElastoDynamics / Adapter /
multigrid.Hierarchy drive it. Checked: exactly ONE JSON line with the contract's keys, the line
survives failing or hanging side measurements (variants, cfg4), and `--impl reference` runs the
oracle. What the numbers are is the GPU run's business; this file only guards the plumbing the
round-end driver depends on."""
import importlib
import io
import json
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeHandle:
    """Scripted device handle: Newton converges in 3 solves, every call is counted."""
    fail_option = None       # (opt, value) whose set_option raises
    hang_lin_step = 0.0      # seconds every lin_step sleeps (watchdog test)
    instances = []

    def __init__(self, problem, device=0, partition=None, comm=None, slab_axis=None):
        self.problem = problem
        self.n_iface_nodes = problem.n_iface_nodes
        self.n_owned = problem.n_dofs
        self.n_cells = problem.mesh.n_cells
        self.calls = {}
        self.options = {}
        self._newton = 0
        self.closed = False
        FakeHandle.instances.append(self)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def close(self):
        self.closed = True

    def set_option(self, opt, value):
        if FakeHandle.fail_option == (opt, value):
            from dealii_adapter_b200 import capi
            raise capi.GraftError(3, "scripted failure")
        self.options[opt] = value

    def mg_attach(self, coarse, child_cells):
        assert child_cells.shape[0] == coarse.n_cells

    def set_traction(self, buf):
        assert len(buf) == self.n_iface_nodes * self.problem.dim
        self._count("set_traction")

    def get_interface_displacement(self):
        self._count("get_interface_displacement")
        return np.zeros(self.n_iface_nodes * self.problem.dim)

    def state_save(self):
        self._count("state_save")

    def state_restore(self):
        self._count("state_restore")

    def nl_begin_step(self):
        self._newton = 0

    def nl_newton_assemble(self):
        self._count("assemble")
        return 10.0 ** (-5 * self._newton)

    def nl_newton_solve(self, type_lin, tol_lin, max_it):
        self._count("solve")
        self._newton += 1
        return 12, 1e-9, 10.0 ** (-4 * self._newton)

    def nl_end_step(self):
        self._count("end_step")

    def lin_assemble_once(self):
        self._count("lin_assemble_once")

    def lin_step(self, type_lin, max_it):
        if FakeHandle.hang_lin_step:
            time.sleep(FakeHandle.hang_lin_step)
        self._count("lin_step")
        return 9, 1e-11

    def nnz(self):
        return 1000

    def spmv_timed(self, which, n):
        return 0.5, 3.0e9

    def profile(self, reset=False):
        from dealii_adapter_b200 import capi
        d = {n: 1.0 for n, _ in capi.GfProfile._fields_}
        d["spmv_launches"] = 100
        d["kernel_launches"] = 1234
        return d

    def synchronize(self):
        pass

    def event_record(self, slot):
        pass

    def event_elapsed_ms(self, a, b):
        return 5.0


@pytest.fixture()
def bench(monkeypatch, native_libs):
    import torch
    from dealii_adapter_b200 import capi
    sys.path.insert(0, ROOT)
    mod = importlib.import_module("bench")
    FakeHandle.fail_option, FakeHandle.hang_lin_step, FakeHandle.instances = None, 0.0, []
    monkeypatch.setattr(capi, "Handle", FakeHandle)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(mod, "CFG4_REPS", (4, 8, 4))
    monkeypatch.setattr(mod, "CPU_SAMPLE_REPS", {"reference": (1, 3, 1), "baseline": (1, 3, 1)})
    monkeypatch.setattr(mod.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(mod.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None,
                                                                "reasons": [], "samples": 0})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "GF_PROFILE_RUN"):
        monkeypatch.delenv(k, raising=False)
    return mod


def run_main(mod, monkeypatch, argv):
    """bench.main() with fd 1 captured (bench.py writes its line with os.write on a dup of fd 1)."""
    r, w = os.pipe()
    saved1 = os.dup(1)
    os.dup2(w, 1)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    try:
        mod.main()
    finally:
        os.dup2(saved1, 1)
        os.close(saved1)
        os.close(w)
        if mod._REAL_STDOUT is not None:
            os.close(mod._REAL_STDOUT)
            mod._REAL_STDOUT = None
    with os.fdopen(r) as f:
        out = f.read()
    lines = [x for x in out.split("\n") if x.strip()]
    assert len(lines) == 1, out
    return json.loads(lines[0])


CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                 "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
                 "e2e", "gpu_launches", "roofline")


def test_one_line_with_the_contract_keys(bench, monkeypatch):
    line = run_main(bench, monkeypatch, ["--reps", "4,8,4", "--steps", "2", "--warmup", "1"])
    for k in CONTRACT_KEYS:
        assert k in line, k
    assert line["metric"] == "newton_step_dofs_per_s" and line["n_gpus"] == 1 and line["warmup"] == 3
    assert line["config"]["newton_solves_in_timed_region"] == 6          # 2 steps x 3 solves
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["newton_solves"] == 6
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert line["roofline_assembly"]["bound"] == "fp64" and line["roofline_assembly"]["frac"] > 0
    assert set(line["variants"]) >= {"matrix_free_operator", "vcycle_fp32_matrices",
                                     "vcycle_all_fp32_operator", "direct_solver_stand_in"}
    assert line["strong_scaling"]["scaling"] == "strong" and line["strong_scaling"]["value"] > 0
    assert all(h.closed for h in FakeHandle.instances)
    fine = FakeHandle.instances[0]
    # implicit coupling k = 2: checkpoint written at the first pass of a window, read after it
    assert fine.calls["state_save"] == fine.calls["state_restore"]
    assert fine.calls["assemble"] == fine.calls["solve"] + fine.calls["end_step"]


def test_failing_variant_is_recorded_and_the_line_survives(bench, monkeypatch):
    from dealii_adapter_b200 import capi
    FakeHandle.fail_option = (capi.OPT_MG_MATRIX_PRECISION, 2)
    line = run_main(bench, monkeypatch, ["--reps", "4,8,4", "--steps", "1", "--no-cpu-baseline"])
    assert "scripted failure" in line["variants"]["error"]
    assert "matrix_free_operator" in line["variants"] and line["value"] > 0
    assert line["strong_scaling"]["value"] > 0


def test_hanging_strong_scaling_part_does_not_cost_the_line(bench, monkeypatch):
    """The watchdog emits the line it has and leaves the process: run in a child process."""
    import subprocess
    code = (
        "import sys, os\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import torch, bench, test_bench_flow as t\n"
        "from dealii_adapter_b200 import capi\n"
        "capi.Handle = t.FakeHandle\n"
        "torch.cuda.is_available = lambda: True\n"
        "torch.cuda.set_device = lambda d: None\n"
        "bench.CFG4_REPS = (4, 8, 4)\n"
        "bench.STRONG_TIMEOUT_S = 2\n"
        "bench.ClockSampler.start = lambda self: None\n"
        "bench.ClockSampler.stop = lambda self: {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}\n"
        "t.FakeHandle.hang_lin_step = 30.0\n"
        "sys.argv = ['bench.py', '--reps', '4,8,4', '--steps', '1', '--no-cpu-baseline', '--no-variants']\n"
        "bench.main()\n" % (ROOT, os.path.join(ROOT, "tests")))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                         env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [x for x in res.stdout.split("\n") if x.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line["value"] > 0 and "did not finish" in line["strong_scaling"]["error"]


def test_reference_arm_line(bench, monkeypatch):
    line = run_main(bench, monkeypatch, ["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
