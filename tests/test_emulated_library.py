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
kernels (TMA SpMV, the mbarrier-pipelined neo-Hookean kernels), the matrix-free operator, the
multigrid and the communicators do not exist in it (GF_ERR_UNSUPPORTED) and stay GPU-tested only:
every (dim, degree) takes the generic cell kernels here, every SpMV the LDG kernel.

Default: a subset that runs in about two minutes. GF_EMU_FULL=1: every body of both GPU files that
needs a serial handle only (about 35 minutes); the files covered: test_gpu_parity,
test_zz_gpu_high_degree, test_gpu_zz_output, test_gpu_zz_reference_pins."""
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
# what each GPU file's own `libs` fixture hands to its tests, out of (capi, solvers, oracle)
LIBS_SHAPE = {HD: lambda c, s, o: (c, s, o), GP: lambda c, s, o: (c, s, o),
              OUT: lambda c, s, o: (c, o), PINS: lambda c, s, o: c}
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
    (HD, "test_hanging_nodes_linear_steps", dict(degree=1)),
    (HD, "test_hanging_nodes_linear_steps", dict(degree=2)),
    (HD, "test_hanging_nodes_nonlinear_step_and_refusals", dict(degree=1)),
    (HD, "test_hanging_nodes_nonlinear_step_and_refusals", dict(degree=2)),
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
]
CASES = (_all_cases(HD) + _all_cases(GP) + _all_cases(OUT) + _all_cases(PINS)) if FULL else FAST


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


def test_hardware_only_parts_are_refused_not_faked(emu_libs):
    """The emulation build must not pretend: multigrid, matrix-free operator and communicators
    answer GF_ERR_UNSUPPORTED, and the binding is restored to the product library afterwards."""
    capi, solvers, orc = emu_libs
    import numpy as np
    from helpers import nl_params
    from dealii_adapter_b200.problem import make_problem
    prob = make_problem(nl_params(poly_degree=2), 3, reps=[2, 2, 2])
    h = capi.Handle(prob)
    with pytest.raises(capi.GraftError) as e:
        h.set_option(capi.OPT_OPERATOR, 1)
        h.nl_newton_assemble()
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    h.set_option(capi.OPT_OPERATOR, 0)
    coarse = capi.Handle(make_problem(nl_params(poly_degree=2), 3, reps=[1, 1, 1]))
    with pytest.raises(capi.GraftError) as e:
        h.mg_attach(coarse, np.arange(8, dtype=np.int64).reshape(1, 8))
    assert e.value.code == capi.GF_ERR_UNSUPPORTED
    with pytest.raises(Exception):
        capi.Comm.unique_id()
    coarse.close()
    h.close()


def test_binding_is_back_on_the_product_library():
    """Outside the fixture capi must not keep the emulation build (a GPU test that ran on it would
    prove nothing)."""
    from dealii_adapter_b200 import build, capi
    assert build.LIB_CUDA.endswith("libgraftfem.so")
    assert capi._lib is None or "emu" not in str(getattr(capi._lib, "_name", ""))
