#!/usr/bin/env python
"""bench.py — Newton-step DoFs/s of the structural hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle restatement of the
                                                            reference's CPU path on a bounded sample)

Workload (config.workload): BASELINE.json configs[2] — nonlinear_elasticity, perpendicular flap 3D,
Q2 hexahedra 24x144x24 cells = 2,081,667 DoFs, neo-Hookean + Newmark, implicit coupling with
checkpoint/restore (k=2 sub-iterations per window, FakeParticipant supplies a constant interface
traction), CG rel. tol 1e-6 ("Residual") preconditioned by the geometric multigrid V-cycle over the
refinement hierarchy (3x18x3 -> 24x144x24 cells; `--precond jacobi` selects the plain block-Jacobi
CG instead). It is the configuration the north_star target is quoted on and it fits one GPU.
N>1: weak scaling — the same flap refined to ~2.05 M DoFs per GPU (WEAK_REPS; N=8: 24x288x96 cells,
16.4 M DoFs), slab-partitioned along y; ghost-DoF halo and dot-product all-reduce by the library's
own kernels over NVLink peer windows (NCCL as fallback); small multigrid levels replicated.

A "step" is one pass through the coupling loop body (save/restore checkpoint, read traction, Newton
loop of assemble + CG solve, Newmark updates, write displacement).
  value : DoFs * (Newton linear solves) / time with the traction already resident in HBM
  e2e   : the same through Solid.step() / Adapter with HOST interface buffers (H2D traction, D2H
          displacement every step)
  roofline : dominant kernel = SpMV inside CG; achieved = bytes of the stored block-row format per
          launch / average launch duration (CUDA events on the library stream, live in the timed
          region); peak = MEASURED_PEAKS.json hbm_gbs
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CELLS_PER_GPU = (24, 144, 24)
# weak scaling: the SAME flap (0.1 x 1 x 0.3) refined so that every GPU keeps ~2.05 M DoFs. Rule:
# each doubling of N halves the LONGEST cell edge (N=1 cells 4.2 x 6.9 x 12.5 mm -> z; 4.2 x 6.9 x
# 6.25 -> y; 4.2 x 3.5 x 6.25 -> z; N=8: 4.2 x 3.5 x 3.1 mm, 16.4 M DoFs = BASELINE configs[4] size
# class). CG iteration counts per Newton solve on these meshes (measured on one GPU, profiles/
# r01_bench_n1_weak_mesh_*.json): 16.5, 13.5, 11.5, 10.8 - the meshes get more isotropic.
WEAK_REPS = {1: (24, 144, 24), 2: (24, 144, 48), 4: (24, 288, 48), 8: (24, 288, 96)}
# CPU arm: the oracle on the SAME flap (0.1 x 1 x 0.3, same load, same solver options) at a coarser
# resolution, so that the slenderness - which sets the SSOR-CG iteration counts - is the workload's
# (a short slab of the fine mesh, as in round 1, is far better conditioned and flatters the CPU).
# Sizes bounded for the run time: ~4 s (reference arm, K steps) / ~25 s (cpu_baseline, 1 step) per
# coupled step. The full-size cfg3 CPU timestep (hours) is a one-off: tools/cpu_cfg3_timestep.py ->
# profiles/r02_cpu_cfg3_full_timestep.jsonl, quoted in the line when present.
CPU_SAMPLE_REPS = {"reference": (4, 24, 4), "baseline": (6, 36, 6)}
TRACTION = (2000.0, 0.0, 0.0)
N_SUB = 2
CFG4_REPS = (128, 1024, 128)   # BASELINE configs[3]: linear Q1 cantilever, 51,171,075 DoFs
STRONG_TIMEOUT_S = 480         # watchdog of the cfg4 part (normally ~40 s incl. set-up)
SIDE_TIMEOUT_S = 420           # watchdog of variants + cpu_baseline (normally ~60 s)


def params():
    from dealii_adapter_b200.problem import SolverParameters
    return SolverParameters(model="neo-Hookean", type_lin="CG", poly_degree=2, scenario="PF",
                            delta_t=0.01, mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6,
                            max_iterations_lin=1.0, max_iterations_NR=10, tol_f=1e-9, tol_u=1e-6,
                            end_time=1e9)


def make_flap(n_layers_y, numbering="lexicographic"):
    """PF flap with cells of the cfg3 size (0.1/24 x 1/144 x 0.3/24) and n_layers_y cell layers."""
    from dealii_adapter_b200.problem import make_problem
    p = params()
    box = ([-0.05, 0.0, 0.0], [0.05, n_layers_y / 144.0, 0.3])
    return make_problem(p, 3, reps=[CELLS_PER_GPU[0], n_layers_y, CELLS_PER_GPU[2]],
                        numbering=numbering, box=box)


def make_flap_reps(reps, numbering="lexicographic"):
    """The perpendicular flap 0.1 x 1 x 0.3 (nonlinear_elasticity.cc:215-219) with `reps` cells."""
    from dealii_adapter_b200.problem import make_problem
    return make_problem(params(), 3, reps=list(reps), numbering=numbering,
                        box=([-0.05, 0.0, 0.0], [0.05, 1.0, 0.3]))


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def full_size_cpu_anchor():
    """The one-off full-size cfg3 CPU timestep, if its log is in profiles/ (same config as `value`)."""
    path = os.path.join(ROOT, "profiles", "r02_cpu_cfg3_full_timestep.jsonl")
    if not os.path.exists(path):
        return None
    rows = [json.loads(x) for x in open(path) if x.strip()]
    solves = [r for r in rows if r.get("event") == "newton_solve"]
    if not solves:
        return None
    done = [r for r in rows if r.get("event") == "done"]
    last = solves[-1]
    return {"what": "oracle on the FULL cfg3 mesh (24x144x24 Q2 cells, 2,081,667 DoFs), offline, "
                    "tools/cpu_cfg3_timestep.py; %s" % ("complete timestep" if done else
                                                        "first %d Newton solve(s)" % len(solves)),
            "newton_solves": len(solves), "cg_iterations": [r["cg_iterations"] for r in solves],
            "seconds": done[0]["seconds"] if done else last["elapsed_s"],
            "dofs_per_s": done[0]["newton_step_dofs_per_s"] if done else last["dofs_per_s_so_far"],
            "threads": rows[0].get("threads"), "cpu": rows[0].get("cpu")}


def cpu_run(n_steps, n_warmup, which="reference"):
    """Oracle (CPU restatement of the reference's path: WorkStream-like threaded assembly, serial
    SSOR-CG as in deal.II) on the bounded sample; returns DoFs/s and details."""
    from oracle import oracle_py as orc
    reps = CPU_SAMPLE_REPS[which]
    prob = make_flap_reps(reps, numbering="cellwise")
    o = orc.Oracle(prob)
    buf = np.tile(TRACTION, prob.n_iface_nodes)
    solves, t_total, cg = 0, 0.0, []
    for s in range(n_warmup + n_steps):
        o.format_precice_to_deal(buf, orc.NL_EXTERNAL_STRESS)
        if s % N_SUB == 0:
            o.save_state()
        t0 = time.perf_counter()
        n, hist = o.nl_timestep()
        dt = time.perf_counter() - t0
        o.format_deal_to_precice(orc.NL_TOTAL_DISPLACEMENT)
        if s % N_SUB != N_SUB - 1:
            o.reload_state()
        if s >= n_warmup:
            solves += n
            t_total += dt
            cg += [int(r[0]) for r in hist]
    out = {"value": prob.n_dofs * solves / t_total, "unit": "DoFs/s", "cores": o.n_threads,
           "kind": "port",
           "sample": "oracle (C++ restatement; the reference cannot be built here: deal.II/preCICE "
                     "absent) on the SAME flap 0.1x1x0.3 with the same load and solver options at "
                     "%dx%dx%d Q2 cells (%d DoFs): %d step(s), %d Newton solves, SSOR-CG iterations "
                     "per solve %d..%d (mean %.0f), CG serial as in deal.II + assembly on %d threads"
                     % (reps[0], reps[1], reps[2], prob.n_dofs, n_steps, solves, min(cg), max(cg),
                        float(np.mean(cg)), o.n_threads),
           "seconds": t_total, "n_dofs": prob.n_dofs, "newton_solves": solves,
           "cg_iterations_per_solve_mean": float(np.mean(cg))}
    anchor = full_size_cpu_anchor()
    if anchor:
        out["full_size_anchor"] = anchor
    return out


def run_cfg4(args, world, rank, local_rank, comm, dist, torch):
    """BASELINE configs[3] / north_star's multi-GPU target: linear_elasticity 3D cantilever, Q1,
    128x1024x128 cells = 51,171,075 DoFs, one-step-theta, STRONG scaling over the ranks (the whole
    problem on one B200 at N = 1: K + A + M resident). One step = assemble_rhs + CG (absolute
    tolerance 1e-10, linear_elasticity.cc:542) + update_displacement. Returns the dict emitted as
    `strong_scaling`."""
    from dealii_adapter_b200 import capi, multigrid, solvers
    from dealii_adapter_b200.problem import SolverParameters, make_problem
    reps = list(CFG4_REPS)
    p = SolverParameters(model="linear", type_lin="CG", poly_degree=1, scenario="PF", delta_t=0.005,
                         mu=0.5e6, nu=0.4, rho=1000.0, theta=0.5, max_iterations_lin=1.0, end_time=1e9)
    t0 = time.perf_counter()
    prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
    H = multigrid.Hierarchy(prob, device=local_rank, world=world, rank=rank, comm=comm, axis=1)
    h = H.fine
    if args.spmv_kernel:
        h.set_option(capi.OPT_SPMV_KERNEL, args.spmv_kernel)
    if args.mg_precision:
        h.set_option(capi.OPT_MG_MATRIX_PRECISION, args.mg_precision)
    t_setup = time.perf_counter() - t0
    buf = np.tile([200.0, 0.0, 0.0], h.n_iface_nodes)
    fp = solvers.FakeParticipant(3, 10 ** 9, p.delta_t, lambda t, it: buf)
    ed = solvers.ElastoDynamics(prob, fp, handle=h)
    ed.adapter.n_interface_nodes = h.n_iface_nodes
    ed.adapter.interface_nodes_ids = np.arange(h.n_iface_nodes, dtype=np.int32)
    h.synchronize()
    t1 = time.perf_counter()
    h.lin_assemble_once()
    h.synchronize()
    t_asm = time.perf_counter() - t1

    def barrier():
        h.synchronize()
        if world > 1:
            dist.barrier()
        h.synchronize()

    steps = max(1, min(args.steps, 10))
    for k in range(max(2, min(args.warmup, 3))):
        ed.step()
    h.set_option(capi.OPT_PROFILE, 2)
    h.profile(reset=True)
    n0 = len(ed.history)
    barrier()
    h.event_record(0)
    t0 = time.perf_counter()
    for k in range(steps):
        ed.step()
    h.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = h.event_elapsed_ms(0, 1)
    prof = h.profile(reset=True)
    h.set_option(capi.OPT_PROFILE, 0)
    t = max(wall, 1e-3 * dev_ms)
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    out = None
    # every rank makes the same library calls in the same order (the call is rank-local today, but
    # a rank-0-only call that completed a deferred collective step hung the 8-GPU run of round 2)
    ms_spmv, nbytes = h.spmv_timed(capi.MAT_SYSTEM, 5)
    if rank == 0:
        peak, src = measured_peak()
        avg_ms = prof["spmv_ms"] / max(1, prof["spmv_launches"])
        its = [int(r[0]) for r in ed.history[n0:]]
        out = {"what": "BASELINE configs[3]: linear_elasticity 3D cantilever Q1 %s cells, %d DoFs, "
                       "one-step-theta, CG abs tol 1e-10 + geometric multigrid; the SAME problem on "
                       "every N (strong scaling), slab-partitioned along y" %
                       ("x".join(map(str, reps)), prob.n_dofs),
               "metric": "timestep_dofs_per_s", "value": prob.n_dofs * steps / t, "unit": "DoFs/s",
               "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": 1e3 * t / steps,
               "n_dofs": prob.n_dofs, "nnz_scalar_per_gpu": h.nnz(), "cg_iterations": its,
               "ms_per_cg_iteration": 1e3 * t / max(1, sum(its)),
               "multigrid_levels": [q.mesh.reps for q in H.problems],
               "multigrid_levels_replicated": H.replicated,
               "setup_s": t_setup, "assemble_once_s": t_asm,
               "roofline": {"bound": "hbm", "kernel": "finest-level SpMV (system matrix A = M + "
                                                      "theta^2 dt^2 K, Q1 rows of 81 values)",
                            "achieved": nbytes / (avg_ms * 1e-3) / 1e9 if avg_ms else None,
                            "peak": peak, "unit": "GB/s", "peak_source": src,
                            "frac": (nbytes / (avg_ms * 1e-3) / 1e9 / peak) if avg_ms else None,
                            "bytes_per_launch": nbytes, "avg_launch_ms": avg_ms,
                            "launches": int(prof["spmv_launches"]), "standalone_launch_ms": ms_spmv,
                            "share_of_step": prof["spmv_ms"] / (1e3 * t)},
               "gpu_launches": int(prof["kernel_launches"])}
    H.close()
    return out


_REAL_STDOUT = None


def arm_watchdog(seconds, on_timeout):
    """The side measurements after the main regions (cfg4 strong scaling) must never cost the
    main line: if they have not finished after `seconds`, rank 0 emits the line it has and every
    rank leaves. Returns the function that disarms it."""
    done = threading.Event()

    def run():
        if not done.wait(seconds):
            try:
                on_timeout()
            finally:
                os._exit(0)
    threading.Thread(target=run, daemon=True).start()
    return done.set


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's
    version banner, torchrun notices) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT, N_SUB
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=CELLS_PER_GPU[1],
                    help="cell layers per GPU along the flap (debug; default = the named workload)")
    ap.add_argument("--reps", default=None,
                    help="a,b,c: cells of the whole flap 0.1 x 1 x 0.3 (debug: run a multi-GPU "
                         "weak-scaling mesh on fewer GPUs)")
    ap.add_argument("--n-sub", type=int, default=N_SUB,
                    help="coupling sub-iterations per time window (checkpoint at the first, restore "
                         "after every non-final one); cfg5 of BASELINE.json: --reps 48,288,60 "
                         "--n-sub 10 on 8 GPUs (tools/bench_cfg5.sh)")
    ap.add_argument("--spmv-kernel", type=int, default=0,
                    help="GF_OPT_SPMV_KERNEL for the whole run (0 = library default; 3 / 6 put the "
                         "fused-dot launches on the 16-consumer-warp kernels as well)")
    ap.add_argument("--mg-precision", type=int, default=0,
                    help="GF_OPT_MG_MATRIX_PRECISION for the whole run (0 FP64 level matrices in the "
                         "V-cycle, 1 FP32 copies, 2 all-FP32 operator)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the cfg4 strong-scaling measurement (51 M-DoF linear Q1 cantilever) "
                         "that follows the main regions")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the matrix-free-operator variant measured after the main regions")
    ap.add_argument("--precond", default="mg", choices=["mg", "jacobi"],
                    help="CG preconditioner: geometric multigrid V-cycle (default) or block-Jacobi")
    args = ap.parse_args()
    N_SUB = max(1, args.n_sub)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.reps:
        reps = tuple(int(x) for x in args.reps.split(","))
    elif args.layers != CELLS_PER_GPU[1] or world not in WEAK_REPS:
        reps = None      # flap of the cfg3 cell size, world * layers cell layers long
    else:
        reps = WEAK_REPS[world]
    workload = ("cfg3 nonlinear_elasticity PF 3D Q2 neo-Hookean, %s cells on %d GPU(s) "
                "(2,081,667 DoFs at N=1; weak scaling: the same flap refined, ~2.05M DoFs per GPU), "
                "implicit coupling k=%d with checkpoint/restore, CG rel tol 1e-6 + %s"
                % ("x".join(map(str, reps)) if reps else "24x%dx24" % (args.layers * world), world,
                   N_SUB,
                   "geometric multigrid V-cycle (Chebyshev/block-Jacobi smoothers)"
                   if args.precond == "mg" else "block-Jacobi"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_run(max(1, args.steps), max(0, min(args.warmup, 1)), "reference")
        line = {"impl": "reference", "metric": "newton_step_dofs_per_s", "value": r["value"],
                "unit": "DoFs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * r["seconds"] / max(1, args.steps), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "cpu_sample": r["sample"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                   "full_size_anchor") if k in r},
                "e2e": {"value": r["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from dealii_adapter_b200 import capi, multigrid, solvers
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, local_rank)

    prob = make_flap_reps(reps) if reps else make_flap(args.layers * world)
    hierarchy = None
    if args.precond == "mg":
        hierarchy = multigrid.Hierarchy(prob, device=local_rank, world=world, rank=rank, comm=comm,
                                        axis=1)
        h = hierarchy.fine
    else:
        part = prob.mesh.partition(1, world, rank) if world > 1 else None
        h = capi.Handle(prob, device=local_rank, partition=part, comm=comm)
    if args.spmv_kernel:
        h.set_option(capi.OPT_SPMV_KERNEL, args.spmv_kernel)
    if args.mg_precision and args.precond == "mg":
        h.set_option(capi.OPT_MG_MATRIX_PRECISION, args.mg_precision)
    n_if = h.n_iface_nodes
    buf = np.tile(TRACTION, n_if)
    participant = solvers.FakeParticipant(3, 10 ** 9, prob.params.delta_t, lambda t, it: buf, N_SUB)
    solid = solvers.Solid(prob, participant, handle=h)
    solid.adapter.n_interface_nodes = n_if
    solid.adapter.interface_nodes_ids = np.arange(n_if, dtype=np.int32)   # Adapter::initialize
    n_dofs_global = prob.n_dofs

    def barrier():
        h.synchronize()
        if world > 1:
            dist.barrier()
        h.synchronize()

    def resident_pass(k):
        # the same loop body with the traction already resident in HBM (no host buffers)
        if k % N_SUB == 0:
            h.state_save()
        h.nl_begin_step()
        solid.solve_nonlinear_timestep()
        h.nl_end_step()
        if k % N_SUB != N_SUB - 1:
            h.state_restore()

    # >= 3 warm-up passes always, except under a profiler (GF_PROFILE_RUN=1: numbers not reported)
    n_warm = args.warmup if os.environ.get("GF_PROFILE_RUN") == "1" else max(3, args.warmup)
    for k in range(n_warm):
        solid.step()

    def cg_iterations(first_step):
        return int(sum(r[0] for rows in solid.history[first_step:] for r in rows))

    sampler = ClockSampler(local_rank)
    # ---- timed region 1: device-resident inputs ("value"). CUDA events bracket every finest-level
    # SpMV launch live in this region (GF_OPT_PROFILE = 2: two events per SpMV launch only; the
    # launch counters of all kernels always run) --------------------------------------------------
    h.set_option(capi.OPT_PROFILE, 2)
    h.profile(reset=True)
    s0 = solid.newton_solves
    h0 = len(solid.history)
    barrier()
    sampler.start()
    h.event_record(0)
    t0 = time.perf_counter()
    for k in range(args.steps):
        resident_pass(k)
    h.event_record(1)
    barrier()
    wall_value = time.perf_counter() - t0
    dev_ms = h.event_elapsed_ms(0, 1)
    prof = h.profile(reset=True)
    solves_value = solid.newton_solves - s0
    cg_its_value = cg_iterations(h0)
    h.set_option(capi.OPT_PROFILE, 0)
    # ---- timed region 2: through the public API with host buffers ("e2e") ---------------------
    s0 = solid.newton_solves
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        solid.step()
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    solves_e2e = solid.newton_solves - s0
    prof_e2e = h.profile(reset=True)
    # ---- diagnostic pass (outside both timed regions): every kernel class bracketed with events
    h.set_option(capi.OPT_PROFILE, 1)
    h.profile(reset=True)
    s0 = solid.newton_solves
    for k in range(N_SUB):
        resident_pass(k)
    barrier()
    prof_full = h.profile(reset=True)
    solves_full = solid.newton_solves - s0
    h.set_option(capi.OPT_PROFILE, 0)
    comm_info = None
    if world > 1:
        kind, n_halo, n_ar = comm.transport()
        halo_us, ar_us = h.comm_timed(50)
        comm_info = {"transport": kind, "halo_exchange_us": halo_us, "allreduce_us": ar_us,
                     "halo_exchanges_issued": n_halo, "allreduces_issued": n_ar,
                     "halo_ms_per_newton_solve_diag": prof_full["halo_ms"] / max(1, solves_full)}

    t_value = max(wall_value, dev_ms * 1e-3)
    if world > 1:
        t = torch.tensor([t_value, wall_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_value, wall_e2e = float(t[0]), float(t[1])
    ms_spmv, spmv_bytes = h.spmv_timed(capi.MAT_TANGENT, 5)   # on every rank: see run_cfg4
    if rank == 0:
        peak, peak_src = measured_peak()
        spmv_avg_ms = prof["spmv_ms"] / max(1, prof["spmv_launches"])
        achieved = spmv_bytes / (spmv_avg_ms * 1e-3) / 1e9
        # DRAM bytes per launch of the SHIPPED default kernel from this round's `ncu --set full`
        # capture of the same matrix (profiles/r02_spmv_ncu_summary.md); null if not captured
        traffic_path = os.path.join(ROOT, "profiles", "r02_spmv_traffic.json")
        traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch") if os.path.exists(traffic_path) else None
        line = {
            "metric": "newton_step_dofs_per_s", "value": n_dofs_global * solves_value / t_value,
            "unit": "DoFs/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": 1e3 * t_value / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_dofs": n_dofs_global, "n_dofs_per_gpu": h.n_owned,
                       "nnz_scalar": h.nnz(), "newton_solves_in_timed_region": solves_value,
                       "cg_iterations_in_timed_region": cg_its_value,
                       "ms_per_cg_iteration": 1e3 * t_value / max(1, cg_its_value),
                       "cg_iterations_per_newton_solve": cg_its_value / max(1, solves_value),
                       "preconditioner": args.precond,
                       "spmv_kernel_option": args.spmv_kernel, "mg_matrix_precision": args.mg_precision,
                       "multigrid_levels": [q.mesh.reps for q in hierarchy.problems] if hierarchy else None,
                       "multigrid_levels_replicated": hierarchy.replicated if hierarchy else None,
                       "l2_policy": "inputs larger than L2 (matrix %.2f GB per GPU streamed every "
                                    "CG iteration)" % (spmv_bytes / 1e9),
                       "device_ms": dev_ms, "parallelism": "slab%d" % world},
            "clocks": clocks,
            "e2e": {"value": n_dofs_global * solves_e2e / wall_e2e, "unit": "DoFs/s",
                    "h2d_bytes_per_step": int(buf.nbytes), "d2h_bytes_per_step": int(buf.nbytes),
                    "ms_per_step": 1e3 * wall_e2e / args.steps, "newton_solves": solves_e2e},
            "gpu_launches": int(prof["kernel_launches"]),
            "roofline": {"bound": "hbm",
                         "kernel": "finest-level SpMV launches of the timed region: spmv_tma2_kernel<3, "
                                   "..., 8, 2, 16, TR> (two TMA rings, 8 gather + 16 consumer warps, "
                                   "transposed row reduction) for the CG vmult and the Chebyshev-"
                                   "smoother / residual vmults of the V-cycle",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "bytes_per_launch": spmv_bytes, "avg_launch_ms": spmv_avg_ms,
                         "launches": int(prof["spmv_launches"]),
                         "standalone_launch_ms": ms_spmv,
                         "share_of_step": prof["spmv_ms"] / (1e3 * t_value)},
            "phase_ms_per_newton_solve": dict(
                {k: v / max(1, solves_full) for k, v in prof_full.items() if k.endswith("_ms")},
                note="diagnostic pass of %d steps outside the timed regions, every kernel class "
                     "bracketed with CUDA events (adds launch gaps to the small kernels)" % N_SUB),
            "variants": {},
        }
        # second roofline: the cell kernel K1 against the MEASURED FP64 (DFMA) peak. Algorithmic
        # flops = what the reference's j <= i loop needs after the pre-contraction T_a = B_a^T D:
        # lower node-block pairs x (27 + 3) FMA x q-points; duration = average K1
        # launch of the diagnostic pass (events around every launch)
        fp64_path = os.path.join(ROOT, "profiles", "fp64_peak.json")
        n_k1 = int(prof_full.get("assemble_cells_launches", 0))
        if prob.dim == 3 and n_k1 > 0 and os.path.exists(fp64_path):
            npc, nq = (prob.degree + 1) ** 3, (prob.degree + 2) ** 3
            flops = 2.0 * 30.0 * (npc * (npc + 1) // 2) * nq * h.n_cells
            k1_ms = prof_full["assemble_cells_ms"] / n_k1
            dfma = float(json.load(open(fp64_path))["dfma_tflops"])
            line["roofline_assembly"] = {
                "bound": "fp64", "kernel": "nl_cells_kernel<3, %d> (K1: element tangents + residuals)"
                                           % prob.degree,
                "achieved": flops / (k1_ms * 1e-3) / 1e12 if k1_ms else None, "peak": dfma,
                "unit": "TFLOP/s", "frac": (flops / (k1_ms * 1e-3) / 1e12 / dfma) if k1_ms else None,
                "peak_source": "measured DFMA peak (profiles/fp64_peak.json, tools/fp64_peak.cu)",
                "flops_per_launch": flops, "avg_launch_ms": k1_ms, "launches": n_k1,
                "cells_per_launch": int(h.n_cells)}
        if comm_info:
            line["comm"] = comm_info
    # ---- from here on the main line exists; everything below is a side measurement that must never
    # cost it: exceptions are recorded in the line, and a watchdog emits the line as it stands if
    # the side measurements hang (rank 0 prints, every rank leaves)
    variants = {}
    snapshot = None
    if rank == 0:
        line["variants"] = variants
        snapshot = json.dumps(line)

    emitted = []

    def emit_line(note=None):
        if rank != 0 or emitted:
            return
        emitted.append(True)
        for attempt in range(3):          # the main thread may be mutating `line` (watchdog call)
            try:
                if note:
                    line["side_measurements"] = note
                emit(line)
                return
            except Exception:
                time.sleep(0.05)
        fallback = json.loads(snapshot)
        fallback["side_measurements"] = note or "line emitted from the snapshot taken after the main regions"
        emit(fallback)

    disarm = arm_watchdog(SIDE_TIMEOUT_S, lambda: emit_line(
        "variants / cpu_baseline did not finish within %d s" % SIDE_TIMEOUT_S))
    # ---- variants benchmarked alongside (north_star: matrix-free operator beside the SpMV) ------
    if not args.no_variants and world == 1 and os.environ.get("GF_PROFILE_RUN") != "1":
        try:
            h.set_option(capi.OPT_OPERATOR, 1)
            for k in range(N_SUB):
                resident_pass(k)
            s0 = solid.newton_solves
            h0v = len(solid.history)
            barrier()
            h.event_record(2)
            t0 = time.perf_counter()
            for k in range(N_SUB):
                resident_pass(k)
            h.event_record(3)
            barrier()
            tv = max(time.perf_counter() - t0, 1e-3 * h.event_elapsed_ms(2, 3))
            ms_mf, mf_bytes = h.spmv_timed(capi.MAT_TANGENT, 5)
            variants["matrix_free_operator"] = {
                "what": "sum-factorised matrix-free tangent (K11) on the finest level instead of the "
                        "assembled BSR SpMV; coarser levels assembled; same CG + V-cycle",
                "value": n_dofs_global * (solid.newton_solves - s0) / tv, "unit": "DoFs/s",
                "steps": N_SUB, "newton_solves": solid.newton_solves - s0,
                "cg_iterations": cg_iterations(h0v), "operator_apply_ms": ms_mf,
                "operator_bytes_per_apply": mf_bytes}
            h.set_option(capi.OPT_OPERATOR, 0)
            if args.precond == "mg":
                # the V-cycle streaming FP32 copies of the level matrices (outer CG stays FP64)
                h.set_option(capi.OPT_MG_MATRIX_PRECISION, 1)
                for k in range(N_SUB):
                    resident_pass(k)
                s0 = solid.newton_solves
                h0v = len(solid.history)
                barrier()
                h.event_record(2)
                t0 = time.perf_counter()
                for k in range(N_SUB):
                    resident_pass(k)
                h.event_record(3)
                barrier()
                tv = max(time.perf_counter() - t0, 1e-3 * h.event_elapsed_ms(2, 3))
                ms32, bytes32 = h.spmv_timed(capi.MAT_MG_F32, 5)
                variants["vcycle_fp32_matrices"] = {
                    "what": "GF_OPT_MG_MATRIX_PRECISION = 1: smoother / residual applications inside "
                            "the V-cycle stream FP32 copies of the level matrices (vectors, "
                            "accumulation, the CG operator and its residual test stay FP64)",
                    "value": n_dofs_global * (solid.newton_solves - s0) / tv, "unit": "DoFs/s",
                    "steps": N_SUB, "newton_solves": solid.newton_solves - s0,
                    "cg_iterations": cg_iterations(h0v), "operator_apply_ms": ms32,
                    "operator_bytes_per_apply": bytes32,
                    "operator_gbs": bytes32 / ms32 / 1e6}
                # stand-alone launches of every TMA kernel kind on the FP64 tangent and its FP32 copy
                # (5: single ring; 3: two rings with 8 gather + 16 consumer warps; 6 = default: 3 with
                # the transposed row reduction)
                kinds = {}
                for kind in (5, 3, 6):
                    h.set_option(capi.OPT_SPMV_KERNEL, kind)
                    m64, b64 = h.spmv_timed(capi.MAT_TANGENT, 5)
                    m32, b32 = h.spmv_timed(capi.MAT_MG_F32, 5)
                    kinds[str(kind)] = {"fp64_ms": m64, "fp64_gbs": b64 / m64 / 1e6,
                                        "fp32_copy_ms": m32, "fp32_copy_gbs": b32 / m32 / 1e6}
                h.set_option(capi.OPT_SPMV_KERNEL, args.spmv_kernel)
                variants["spmv_kernel_kinds"] = dict(
                    kinds, what="GF_OPT_SPMV_KERNEL: stand-alone y = A x launches (no fused dot), "
                                "bitwise equal results for all kinds")
                h.set_option(capi.OPT_MG_MATRIX_PRECISION, 2)
                for k in range(N_SUB):
                    resident_pass(k)
                s0 = solid.newton_solves
                h0v = len(solid.history)
                barrier()
                h.event_record(2)
                t0 = time.perf_counter()
                for k in range(N_SUB):
                    resident_pass(k)
                h.event_record(3)
                barrier()
                tv = max(time.perf_counter() - t0, 1e-3 * h.event_elapsed_ms(2, 3))
                ms32, bytes32 = h.spmv_timed(capi.MAT_MG_F32, 5)
                variants["vcycle_all_fp32_operator"] = {
                    "what": "GF_OPT_MG_MATRIX_PRECISION = 2: the V-cycle's operator applications in "
                            "single precision throughout (FP32 matrix copy, x staged and accumulated "
                            "in FP32; vectors in HBM, the CG operator, its residual test and the "
                            "Newton tolerances stay FP64)",
                    "value": n_dofs_global * (solid.newton_solves - s0) / tv, "unit": "DoFs/s",
                    "steps": N_SUB, "newton_solves": solid.newton_solves - s0,
                    "cg_iterations": cg_iterations(h0v), "operator_apply_ms": ms32,
                    "operator_gbs": bytes32 / ms32 / 1e6}
                h.set_option(capi.OPT_MG_MATRIX_PRECISION, args.mg_precision)
                # "Solver type = Direct" (the shipped default, parameters.prm:43): the stand-in is
                # the same CG from a zero guess to 1e-13 relative; its cost with the V-cycle
                solid.parameters.type_lin = "Direct"
                for k in range(N_SUB):
                    resident_pass(k)
                s0 = solid.newton_solves
                barrier()
                h.event_record(2)
                t0 = time.perf_counter()
                for k in range(N_SUB):
                    resident_pass(k)
                h.event_record(3)
                barrier()
                tv = max(time.perf_counter() - t0, 1e-3 * h.event_elapsed_ms(2, 3))
                variants["direct_solver_stand_in"] = {
                    "what": "type_lin = Direct: UMFPACK (nonlinear_elasticity.cc:1192-1200) is replaced "
                            "by the multigrid-preconditioned CG run to 1e-13 relative from a zero guess",
                    "value": n_dofs_global * (solid.newton_solves - s0) / tv, "unit": "DoFs/s",
                    "steps": N_SUB, "newton_solves": solid.newton_solves - s0,
                    "ms_per_newton_solve": 1e3 * tv / max(1, solid.newton_solves - s0)}
                solid.parameters.type_lin = "CG"
        except Exception as exc:      # a failing side measurement must not cost the main line
            variants["error"] = "%s: %s" % (type(exc).__name__, exc)
            for opt, val in ((capi.OPT_OPERATOR, 0), (capi.OPT_SPMV_KERNEL, args.spmv_kernel),
                             (capi.OPT_MG_MATRIX_PRECISION, args.mg_precision)):
                try:
                    h.set_option(opt, val)
                except Exception:
                    pass

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_run(1, 0, "baseline")
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                      "full_size_anchor") if k in r}
        except Exception as exc:
            line["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    disarm()
    try:
        if hierarchy:
            hierarchy.close()
        else:
            h.close()
    except Exception as exc:
        if rank == 0:
            line["close_error"] = "%s: %s" % (type(exc).__name__, exc)
    # ---- north_star's multi-GPU target in the same driver-run line: cfg4, STRONG scaling -------
    if not args.no_strong and os.environ.get("GF_PROFILE_RUN") != "1" and args.precond == "mg":
        def give_up():
            if rank == 0:
                line["strong_scaling"] = {"error": "did not finish within %d s" % STRONG_TIMEOUT_S}
            emit_line()
        disarm = arm_watchdog(STRONG_TIMEOUT_S, give_up)
        try:
            strong = run_cfg4(args, world, rank, local_rank, comm, dist, torch)
        except Exception as exc:      # must not cost the main line
            strong = {"error": "%s: %s" % (type(exc).__name__, exc)}
        disarm()
        if rank == 0:
            line["strong_scaling"] = strong
    emit_line()
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
