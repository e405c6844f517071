"""The library's own sources on a machine WITHOUT a GPU: tests/cuda_emu/make_emu_library.py compiles
the C-ABI, the host glue and every kernel that needs neither TMA nor mbarriers with g++ against a
CUDA stand-in (a CTA = cooperative fibers, warp shuffles among 32 of them; launches synchronous),
and this file drives that build through the SAME ctypes binding and the SAME test bodies the GPU
suite runs (`tests/test_gpu_parity.py`, `tests/test_zz_gpu_high_degree.py`) against the oracle.

What it is for: several parts of the library were written in sessions without GPU access
(polynomial degree >= 3, general cells, the band Cholesky, hanging-node constraint lines). Their
kernels were checked in isolation (tests/test_cuda_emulation.py); this file also executes what sits
between the C-ABI and the kernels - numbering and sparsity pattern, scatter map, reductions, CG,
Newton / theta-scheme entry points, export, output - before the first run on hardware.
What it is NOT: a product path. Nothing outside tests/ can load the emulation build; the tuned
kernels (TMA SpMV, the mbarrier-pipelined neo-Hookean kernels), the single-launch coarsest-level
solver and the NCCL transport do not exist in it and stay GPU-tested only: every (dim, degree)
takes the generic cell kernels here, every SpMV the LDG kernel, the coarsest multigrid level its
multi-launch fallback; ranks are processes whose peer windows are shared-memory files.

Default: a subset that runs in about four minutes (incl. bench.py on two emulated ranks). GF_EMU_FULL=1: every body of both GPU files that
needs a serial handle only (118 tests, about 110 minutes); the files covered: test_gpu_parity,
test_zz_gpu_high_degree, test_gpu_zz_output, test_gpu_zz_reference_pins, test_gpu_multigrid,
test_gpu_matfree - and bench.py itself."""
import importlib
import inspect
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FULL = os.environ.get("GF_EMU_FULL") == "1"


@pytest.fixture(scope="module")
def emu_lib_path(native_libs):
    sys.path.insert(0, os.path.join(HERE, "cuda_emu"))
    import make_emu_library
    return make_emu_library.build()


@pytest.fixture()
def emu_libs(emu_lib_path):
    """(capi, solvers, oracle) with capi bound to the emulation build for the time of ONE test."""
    from dealii_adapter_b200 import build, capi, solvers
    from oracle import oracle_py
    saved = (build.LIB_CUDA, capi._lib, build.build_cuda)
    build.LIB_CUDA, capi._lib = emu_lib_path, None
    build.build_cuda = lambda *a, **k: emu_lib_path   # the GPU files' fixtures rebuild the product
    capi.lib()
    try:
        yield capi, solvers, oracle_py
    finally:
        build.LIB_CUDA, capi._lib, build.build_cuda = saved


def _body(module, name):
    fn = getattr(importlib.import_module(module), name)
    return getattr(fn, "__wrapped__", fn)            # without the first-run bookkeeping wrapper


def _all_cases(module):
    """every (test name, parameters) of a GPU test module that needs nothing but `libs`"""
    mod = importlib.import_module(module)
    out = []
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or fn.__module__ != mod.__name__:
            continue
        sig = inspect.signature(getattr(fn, "__wrapped__", fn)).parameters
        if "libs" not in sig or "tmp_path" in sig or "native_libs" in sig:
            continue
        params = [{}]
        for m in getattr(fn, "pytestmark", []):
            if m.name == "parametrize":
                names = [x.strip() for x in m.args[0].split(",")]
                params = [dict(zip(names, v if len(names) > 1 else (v,))) for v in m.args[1]]
        out += [(module, name, p) for p in params]
    return out


HD, GP = "test_zz_gpu_high_degree", "test_gpu_parity"
OUT, PINS = "test_gpu_zz_output", "test_gpu_zz_reference_pins"
MG, MF, F32 = "test_gpu_multigrid", "test_gpu_matfree", "test_gpu_zz_mg_f32"
# what each GPU file's own `libs` fixture hands to its tests, out of (capi, solvers, oracle)
LIBS_SHAPE = {HD: lambda c, s, o: (c, s, o), GP: lambda c, s, o: (c, s, o),
              OUT: lambda c, s, o: (c, o), PINS: lambda c, s, o: c,
              MG: lambda c, s, o: (c, s, _multigrid(), o), MF: lambda c, s, o: (c, s, _multigrid(), o),
              F32: lambda c, s, o: (c, s, _multigrid())}
# the all-FP32 operator (x staged and accumulated in FP32) exists as a TMA kernel only
HARDWARE_ONLY = {"test_all_fp32_operator_is_a_single_precision_spmv",
                 "test_all_fp32_vcycle_keeps_newton_counts_and_displacements"}


def _multigrid():
    from dealii_adapter_b200 import multigrid
    return multigrid
FAST = [
    # ---- degree >= 3 (generic kernels, FESystem numbering map, pattern with (2p+1)^dim blocks)
    (HD, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=2, degree=3, reps=[3, 4], numbering="cellwise")),
    (HD, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=2, degree=5, reps=[2, 2], numbering="lexicographic")),
    (HD, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=3, degree=3, reps=[1, 2, 2], numbering="lexicographic")),
    (HD, "test_linear_matrices_and_steps_match_oracle", dict(dim=2, degree=4, reps=[2, 4], numbering="lexicographic")),
    (HD, "test_linear_matrices_and_steps_match_oracle", dict(dim=3, degree=3, reps=[2, 2, 1], numbering="component_wise")),
    (HD, "test_output_fields_match_oracle", dict(dim=2, degree=3, reps=[3, 2])),
    (HD, "test_output_fields_match_oracle", dict(dim=3, degree=3, reps=[1, 2, 1])),
    (HD, "test_multigrid_and_matrix_free_are_refused_above_degree_2", {}),
    # ---- the reference's own assembly blocks (golden vectors) through the C-ABI
    (HD, "test_device_cell_assembly_equals_the_reference_assembly_block", dict(case=5)),
    (HD, "test_device_cell_assembly_equals_the_reference_assembly_block", dict(case=8)),
    (HD, "test_device_cell_assembly_equals_the_reference_assembly_block", dict(case=10)),
    (HD, "test_device_linear_stiffness_and_loading_equal_the_reference_loops", dict(case=4)),
    (HD, "test_device_linear_stiffness_and_loading_equal_the_reference_loops", dict(case=6)),
    # ---- general (non-affine) cells
    (HD, "test_distorted_mesh_nonlinear_tangent_residual_and_output", dict(dim=2, degree=3, reps=[3, 3], numbering="cellwise")),
    (HD, "test_distorted_mesh_nonlinear_tangent_residual_and_output", dict(dim=3, degree=2, reps=[2, 3, 2], numbering="cellwise")),
    (HD, "test_distorted_mesh_linear_matrices_and_steps", dict(dim=3, degree=1, reps=[3, 4, 2])),
    (HD, "test_distorted_mesh_linear_matrices_and_steps", dict(dim=2, degree=3, reps=[2, 4])),
    # ---- hanging-node constraint lines
    (HD, "test_hanging_nodes_linear_steps", dict(dim=2, degree=1)),
    (HD, "test_hanging_nodes_linear_steps", dict(dim=2, degree=2)),
    (HD, "test_hanging_nodes_linear_steps", dict(dim=3, degree=1)),
    (HD, "test_hanging_nodes_linear_steps", dict(dim=3, degree=2)),
    (HD, "test_hanging_nodes_nonlinear_step_direct_and_cg", dict(dim=2, degree=2)),
    (HD, "test_hanging_nodes_nonlinear_step_direct_and_cg", dict(dim=3, degree=1)),
    (HD, "test_hanging_nodes_nonlinear_step_direct_and_cg", dict(dim=3, degree=2)),
    # ---- 'Solver type = Direct': band Cholesky behind the Newton / theta-scheme entry points
    #      (smaller meshes than the GPU run of the same bodies)
    (HD, "test_direct_solver_band_cholesky_nonlinear", dict(dim=2, degree=3, scenario="FSI3", reps=[6, 1], load=(0.0, -1500.0))),
    (HD, "test_direct_solver_band_cholesky_nonlinear", dict(dim=3, degree=2, scenario="PF", reps=[1, 3, 1], load=(1500.0, 0.0, 0.0))),
    # ---- the round-1/2 parity bodies: degrees 1 and 2 run the generic kernels here
    (GP, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=2, degree=2, reps=[3, 4], numbering="component_wise")),
    (GP, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=3, degree=1, reps=[2, 3, 2], numbering="cellwise")),
    (GP, "test_nonlinear_tangent_and_residual_match_oracle", dict(dim=3, degree=2, reps=[3, 2, 2], numbering="lexicographic")),
    (GP, "test_chunked_element_buffer_gives_identical_matrix", {}),
    (GP, "test_det_F_nonpositive_is_reported", {}),
    (GP, "test_spmv_and_cg_match_oracle_matrix", dict(dim=2, reps=[2, 6])),
    (GP, "test_linear_matrices_and_steps_match_oracle", dict(dim=3, degree=1, reps=[3, 4, 2], numbering="cellwise")),
    (GP, "test_state_save_restore_and_interface_roundtrip", {}),
    (GP, "test_partitioned_assembly_matches_global_rows", {}),
    (GP, "test_newton_counts_and_watchpoint_match_oracle", dict(dim=2, scenario="FSI3", reps=[6, 1], load=(0.0, -1500.0))),
    (GP, "test_newton_counts_and_watchpoint_match_oracle", dict(dim=3, scenario="PF", reps=[1, 3, 1], load=(1500.0, 0.0, 0.0))),
    # ---- output path (DataOut patches + Postprocessor) and the reference pins of round 1
    (OUT, "test_postprocess_matches_oracle", dict(model="nl", dim=3, degree=2, reps=[2, 3, 2], numbering="lexicographic")),
    (OUT, "test_postprocess_matches_oracle", dict(model="lin", dim=2, degree=1, reps=[5, 7], numbering="cellwise")),
    (OUT, "test_postprocess_rejects_an_inverted_displaced_cell", {}),
    (PINS, "test_device_cell_assembly_equals_the_reference_assembly_block", dict(case=1)),
    (PINS, "test_device_linear_stiffness_and_loading_equal_the_reference_loops", dict(case=2)),
    (PINS, "test_device_newmark_updates_equal_the_reference_members", dict(case=0)),
    (PINS, "test_device_theta_scheme_rhs_equals_the_reference_block", dict(case=1)),
    # ---- geometric multigrid (transfer, Chebyshev smoothers, level operators; the coarsest level
    #      by its multi-launch fallback) and the matrix-free tangent
    (MG, "test_vcycle_is_symmetric_positive_definite", dict(dim=3, degree=2, reps=[4, 8, 4], numbering="lexicographic")),
    (MG, "test_vcycle_is_symmetric_positive_definite", dict(dim=2, degree=2, reps=[8, 16], numbering="component_wise")),
    (MG, "test_vcycle_approximates_the_inverse", {}),
    (MG, "test_attach_rejects_bad_hierarchies", {}),
    (MG, "test_multigrid_linear_model_matches_block_jacobi", {}),
    (MF, "test_matrix_free_operator_matches_assembled_tangent", dict(degree=2, reps=[2, 5, 3], numbering="component_wise")),
    (MF, "test_matrix_free_operator_matches_assembled_tangent", dict(degree=1, reps=[4, 5, 3], numbering="lexicographic")),
    (MF, "test_matrix_free_rejected_where_unsupported", {}),
    # ---- FP32 copies of the level operators inside the V-cycle (GF_OPT_MG_MATRIX_PRECISION = 1)
    (F32, "test_f32_operator_copy_matches_fp64_to_single_precision", dict(dim=3, degree=2, reps=[4, 8, 4], numbering="lexicographic")),
    (F32, "test_f32_vcycle_is_symmetric_positive_definite_and_close_to_fp64", {}),
]
CASES = [c for c in (_all_cases(HD) + _all_cases(GP) + _all_cases(OUT) + _all_cases(PINS) +
                     _all_cases(MG) + _all_cases(MF) + _all_cases(F32))
         if c[1] not in HARDWARE_ONLY] if FULL else FAST


def _id(case):
    module, name, params = case
    return "%s-%s" % (name[5:45], "-".join(str(v).replace(" ", "") for v in params.values()))


@pytest.mark.parametrize("case", CASES, ids=[_id(c) for c in CASES])
def test_gpu_test_body_on_the_emulated_library(emu_libs, monkeypatch, case):
    module, name, params = case
    body = _body(module, name)
    kwargs = dict(params)
    if "monkeypatch" in inspect.signature(body).parameters:
        kwargs["monkeypatch"] = monkeypatch
    body(LIBS_SHAPE[module](*emu_libs), **kwargs)


def test_cpp_host_driver_on_the_emulated_library(emu_lib_path, native_libs, tmp_path, monkeypatch):
    """The C++ drop-in driver (elasticity_2d: Parameters / Time / Adapter / Solid / ElastoDynamics
    mirrors) end to end without a GPU: the emulation build is LD_PRELOADed under the executable,
    so its gf_* calls resolve there. Bodies of the GPU tests: the reference's shipped
    parameters.prm (linear, degree 3, Direct -> band Cholesky) and the neo-Hookean FSI3 run, both
    against the oracle's watch point."""
    monkeypatch.setenv("LD_PRELOAD", emu_lib_path)
    monkeypatch.setenv("GF_DIRECT_SOLVER", "auto")
    exes = native_libs.build_elasticity()
    a, b = tmp_path / "shipped", tmp_path / "nonlinear"
    a.mkdir()
    b.mkdir()
    _body(HD, "test_shipped_parameter_file_runs_unchanged_through_the_cpp_driver")(a, native_libs)
    _body("test_host_driver", "test_coupled_nonlinear_run_matches_oracle_watchpoint")(exes, b, native_libs)


@pytest.mark.skipif(not FULL, reason="GF_EMU_FULL=1 (the default run executes bench.py on two emulated ranks)")
def test_bench_py_end_to_end_on_the_emulated_library(emu_libs, monkeypatch):
    """bench.py itself - Hierarchy, Solid, the timed regions, the stand-alone SpMV timing after a
    deferred tangent, the cfg4 strong-scaling part (and, with GF_EMU_FULL=1, every variant) -
    driving the REAL library source on a tiny flap: the plumbing of the round-end driver run,
    executed. The numbers mean nothing here; that there is exactly one well-formed line does."""
    import io
    import json
    import torch
    bench = importlib.import_module("bench")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(bench, "CFG4_REPS", (2, 4, 2))
    monkeypatch.setattr(bench, "CPU_SAMPLE_REPS", {"reference": (1, 3, 1), "baseline": (1, 3, 1)})
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None,
                                                                 "reasons": [], "samples": 0})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "GF_PROFILE_RUN"):
        monkeypatch.delenv(k, raising=False)
    argv = ["bench.py", "--reps", "2,2,2", "--steps", "1"]
    monkeypatch.setattr(sys, "argv", argv)
    r, w = os.pipe()
    saved1 = os.dup(1)
    os.dup2(w, 1)
    try:
        bench.main()
    finally:
        os.dup2(saved1, 1)
        os.close(saved1)
        os.close(w)
        if bench._REAL_STDOUT is not None:
            os.close(bench._REAL_STDOUT)
            bench._REAL_STDOUT = None
    with os.fdopen(r) as f:
        lines = [x for x in f.read().split("\n") if x.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["metric"] == "newton_step_dofs_per_s" and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["config"]["multigrid_levels"] == [[2, 2, 2], [1, 1, 1]]
    assert line["config"]["newton_solves_in_timed_region"] >= 3
    assert line["gpu_launches"] > 100 and line["roofline"]["launches"] > 10
    assert 0 < line["roofline"]["share_of_step"] < 1
    assert "error" not in line["strong_scaling"] and line["strong_scaling"]["cg_iterations"][0] > 0
    assert "close_error" not in line and "side_measurements" not in line
    if FULL:
        assert "error" not in line["variants"]
        assert set(line["variants"]) >= {"matrix_free_operator", "vcycle_fp32_matrices",
                                         "vcycle_all_fp32_operator", "direct_solver_stand_in"}


MULTIRANK_CASES = ["nl_jacobi", "lin_jacobi", "nl_mg_small", "lin_mg_small",
                   "nl_mg_small_partitioned_coarse", "lin_mg_small_partitioned_coarse",
                   "nl_jacobi_q3", "lin_jacobi_q3"]


@pytest.mark.parametrize("world", [2, 4] if FULL else [2])
def test_partitioned_runs_on_emulated_ranks_reproduce_the_single_rank_run(emu_libs, emu_lib_path,
                                                                          tmp_path, world):
    """The multi-rank path without a GPU: the ranks are PROCESSES (torch.distributed.run, gloo for
    the bootstrap) that each load the emulation build; comm.cu's peer windows - mailboxes, flags,
    the push / wait kernels, the gather of the reduction partials - run as they are on
    shared-memory files behind cudaIpcGetMemHandle / cudaIpcOpenMemHandle. Same worker and same
    assertions as tests/test_gpu_multirank.py on meshes sized for the emulation: Newton and CG
    counts identical and the written interface displacement BITWISE equal to the single-rank run,
    with the coarse multigrid level replicated (vector all-reduce of the restricted residual) and
    slab-partitioned (halos on both levels), through the deferred tangent completion."""
    import mgpu_worker as w
    from test_gpu_multirank import _compare, _spawn
    # default: the linear model with halos on both multigrid levels; the neo-Hookean model, the
    # deferred tangent and the replicated coarse level run in test_bench_py_on_emulated_ranks
    cases = MULTIRANK_CASES if FULL else ["lin_mg_small_partitioned_coarse"]
    ref = {}
    for name in cases:
        hist, written, levels = w.run_case(name, 1, 0, 0, None)
        ref[name] = {"written": written, "history": hist, "levels": levels}
    os.environ["GF_TEST_EMU_LIB"] = emu_lib_path
    os.environ["GF_WORKER_DEADLINE_S"] = "1500"
    os.environ["GF_TEST_P2P_TIMEOUT_S"] = "120"     # CPU ranks on a possibly busy machine
    try:
        got = _spawn(world, "ipc", str(tmp_path / "ranks.pkl"), cases=",".join(cases), timeout=1600)
    finally:
        del os.environ["GF_TEST_EMU_LIB"], os.environ["GF_WORKER_DEADLINE_S"]
        del os.environ["GF_TEST_P2P_TIMEOUT_S"]
    assert got["transport"][0] == "peer_windows" and got["transport"][1] > 0
    for name in cases:
        _compare(ref, got, name, True)
    if "lin_mg_small" in cases:
        assert got["lin_mg_small"]["levels"] == (2, [False, True])          # coarse level replicated
    assert got["lin_mg_small_partitioned_coarse"]["levels"] == (2, [False, False])


@pytest.mark.parametrize("world", [2, 4] if FULL else [2])
def test_bench_py_on_emulated_ranks(emu_lib_path, tmp_path, world):
    """`bench.py --gpus N` as the driver launches it (torch.distributed.run, one process per
    rank), unchanged, on N emulated ranks (tests/bench_emu_worker.py redirects nccl -> gloo and
    the communicator bootstrap -> gf_comm_ipc_*): every library call and every collective of
    every rank is real. A rank-asymmetric call sequence - what hung one 8-GPU run of round 2 -
    would stop here within the peer-window timeout. Checked: exit code 0, ONE line from rank 0,
    weak + strong parts present, the same Newton / CG counts for every N."""
    import json
    import socket
    import subprocess
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, GF_TEST_EMU_LIB=emu_lib_path, GF_P2P_TIMEOUT_S="120", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "bench_emu_worker.py"), "--gpus", str(world), "--reps", "2,8,2",
           "--steps", "1"]
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=900)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(proc.pid, signal.SIGKILL)
        raise AssertionError("bench.py on %d emulated ranks did not finish" % world)
    assert proc.returncode == 0, err[-3000:]
    lines = [x for x in out.split("\n") if x.startswith("{")]
    assert len(lines) == 1, out
    line = json.loads(lines[0])
    assert line["n_gpus"] == world and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["comm"]["transport"] == "peer_windows" and line["comm"]["halo_exchanges_issued"] > 0
    assert line["config"]["multigrid_levels_replicated"] == [False, True]
    # partition independent: the counts of the single-rank run of the same mesh (measured once
    # on the emulation: 4 Newton solves, 58 CG iterations in the one timed pass)
    assert line["config"]["newton_solves_in_timed_region"] == 4
    assert line["config"]["cg_iterations_in_timed_region"] == 58
    strong = line["strong_scaling"]
    assert "error" not in strong and strong["n_gpus"] == world and strong["cg_iterations"] == [11]


def test_a_rank_that_calls_a_collective_alone_gets_an_error_not_a_hang(emu_lib_path, tmp_path):
    """What hung the 8-GPU run of round 2 was a collective call made by one rank only. Since then a
    peer-window wait gives up after GF_P2P_TIMEOUT_S, raises the communicator's (sticky) error flag
    and the next comm_check of the host turns it into GF_ERR_NCCL. Never exercised on hardware;
    here two emulated ranks do exactly that (tests/comm_timeout_worker.py)."""
    import pickle
    import socket
    import subprocess
    from dealii_adapter_b200 import capi
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = str(tmp_path / "report.pkl")
    env = dict(os.environ, GF_TEST_EMU_LIB=emu_lib_path, GF_P2P_TIMEOUT_S="3", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(port), os.path.join(HERE, "comm_timeout_worker.py"), out],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    rep = pickle.load(open(out, "rb"))
    code, msg, seconds = rep["first"]
    assert code == capi.GF_ERR_NCCL and "timed out" in msg and 2.5 <= seconds < 60
    code2, msg2, seconds2 = rep["second"]
    assert code2 == capi.GF_ERR_NCCL and seconds2 < 2.5          # sticky: no second timeout


SCHEDULE_SUBSET = ("tangent or hanging or distorted or output or cell_assembly or vcycle or "
                   "matrix_free or spmv_and_cg or det_F or chunked")


@pytest.mark.parametrize("schedule", ["reverse", "random:11", "random:12"] if FULL else ["reverse"])
def test_kernels_do_not_depend_on_the_thread_schedule(emu_lib_path, schedule):
    """The fibers of a CTA run in thread order between two barriers, which would HIDE a missing
    __syncthreads of the kind 'a lower thread writes, a higher thread reads'. The same tests again
    with the fibers scheduled in reverse / in a new random order every sweep (GF_EMU_SCHEDULE, read
    once per process - hence the child process): same results, including the bitwise
    run-to-run checks inside the bodies."""
    import subprocess
    if os.environ.get("GF_EMU_SCHEDULE"):
        pytest.skip("already running under GF_EMU_SCHEDULE")
    env = dict(os.environ, GF_EMU_SCHEDULE=schedule)
    env.pop("GF_EMU_FULL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(HERE, "test_cuda_emulation.py"),
                        os.path.join(HERE, "test_emulated_library.py"), "-k",
                        "test_cuda_emulation or (test_gpu_test_body and (%s))" % SCHEDULE_SUBSET],
                       env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_hanging_nodes_next_to_a_dirichlet_boundary(emu_libs, dim, degree):
    """The clamp on y = 0 instead of x = 0: hanging nodes ON the clamped face are plain Dirichlet
    dofs after AffineConstraints::close(), and lines of their neighbours lose the masters that are
    Dirichlet dofs (tests/helpers.py mimics close()). Newton step against the by-definition
    reference with the factor-preconditioned CG and with the tight block-Jacobi CG."""
    import numpy as np
    from helpers import hanging_node_problem, nl_params, reference_nonlinear_step, rel_err
    capi, solvers, orc = emu_libs
    p = nl_params(poly_degree=degree, type_lin="Direct", scenario="PF", delta_t=0.01)
    prob = hanging_node_problem(p, degree, dim, clamp_axis=1)
    dof, ptr, master, weight = prob.extra["constraint_lines"]
    assert not prob.constrained[dof].any() and not prob.constrained[master].any()
    assert (np.diff(ptr) < (degree + 1) ** (dim - 1)).any()      # some line lost a master
    buf = np.tile([300.0, -1500.0, 300.0][:dim], prob.n_iface_nodes)
    ref = reference_nonlinear_step(orc, prob, buf)
    for mode in (0, 2):
        part = solvers.FakeParticipant(dim, 1, p.delta_t, lambda t, it: buf)
        solid = solvers.Solid(prob, part)
        solid.handle.set_option(capi.OPT_DIRECT_SOLVER, mode)
        solid.run()
        assert rel_err(solid.handle.get_vector(capi.NL_TOTAL_DISPLACEMENT), ref) < 1e-8
        solid.handle.close()


def test_the_product_library_contains_no_emulation_code(native_libs):
    """GF_CUDA_EMULATION is a define of tests/cuda_emu/make_emu_library.py only: the product's build
    recipe never sets it and libgraftfem.so carries no symbol of the stand-in."""
    import subprocess
    from dealii_adapter_b200 import build
    recipe = open(build.__file__).read()
    assert "GF_CUDA_EMULATION" not in recipe and "cuda_emu" not in recipe
    assert not any("EMULATION" in f for f in build.NVCC_FLAGS)
    lib = native_libs.build_cuda()
    syms = subprocess.run(["nm", "-C", lib], capture_output=True, text=True).stdout
    assert "gf_emu" not in syms and "switch_context" not in syms
    assert "spmv_tma2_kernel" in syms            # the hardware kernels are what it is made of


def test_binding_is_back_on_the_product_library():
    """Outside the fixture capi must not keep the emulation build (a GPU test that ran on it would
    prove nothing)."""
    from dealii_adapter_b200 import build, capi
    assert build.LIB_CUDA.endswith("libgraftfem.so")
    assert capi._lib is None or "emu" not in str(getattr(capi._lib, "_name", ""))
