#!/usr/bin/env python
"""BASELINE.json configs[3]: linear_elasticity 3D cantilever, Q1 hexahedra, Newmark/one-step-theta,
partitioned CG. Default 128 x 1024 x 128 cells = 51,171,075 DoFs (SURVEY 8 size table), STRONG
scaling over the ranks torchrun starts (1 rank: the whole problem on one B200, K + A + M resident).

  python tools/bench_cfg4.py [--reps 128,1024,128] [--steps 4] [--warmup 2] [--precond mg|jacobi]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_cfg4.py

Prints one JSON line: timestep DoFs/s (one linear solve per step: assemble_rhs + CG + update),
SpMV GB/s of the stored block format against the measured HBM peak."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", default="128,1024,128")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precond", default="mg", choices=["mg", "jacobi"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from dealii_adapter_b200 import capi, multigrid, solvers
    from dealii_adapter_b200.problem import SolverParameters, make_problem
    torch.cuda.set_device(lr)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, lr)
    reps = [int(x) for x in args.reps.split(",")]
    p = SolverParameters(model="linear", type_lin="CG", poly_degree=1, scenario="PF", delta_t=0.005,
                         mu=0.5e6, nu=0.4, rho=1000.0, theta=0.5, max_iterations_lin=1.0, end_time=1e9)
    t0 = time.perf_counter()
    prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
    t_mesh = time.perf_counter() - t0
    H = None
    if args.precond == "mg":
        H = multigrid.Hierarchy(prob, device=lr, world=world, rank=rank, comm=comm, axis=1)
        h = H.fine
    else:
        part = prob.mesh.partition(1, world, rank) if world > 1 else None
        h = capi.Handle(prob, device=lr, partition=part, comm=comm)
    t_create = time.perf_counter() - t0 - t_mesh
    buf = np.tile([200.0, 0.0, 0.0], h.n_iface_nodes)
    fp = solvers.FakeParticipant(3, 10 ** 9, p.delta_t, lambda t, it: buf)
    ed = solvers.ElastoDynamics(prob, fp, handle=h)
    ed.adapter.n_interface_nodes = h.n_iface_nodes
    ed.adapter.interface_nodes_ids = np.arange(h.n_iface_nodes, dtype=np.int32)
    h.synchronize(); t1 = time.perf_counter()
    h.lin_assemble_once()
    h.synchronize(); t_asm = time.perf_counter() - t1

    def barrier():
        h.synchronize()
        if world > 1:
            dist.barrier()
        h.synchronize()

    for k in range(args.warmup):
        ed.step()
    h.set_option(capi.OPT_PROFILE, 2)
    h.profile(reset=True)
    n0 = len(ed.history)
    barrier()
    h.event_record(0)
    t0 = time.perf_counter()
    for k in range(args.steps):
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
    if rank == 0:
        peak, src = bench.measured_peak()
        ms_spmv, nbytes = h.spmv_timed(capi.MAT_SYSTEM, 5)
        avg_ms = prof["spmv_ms"] / max(1, prof["spmv_launches"])
        its = [int(r[0]) for r in ed.history[n0:]]
        line = {"metric": "timestep_dofs_per_s", "value": prob.n_dofs * args.steps / t, "unit": "DoFs/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cfg4 linear_elasticity 3D cantilever Q1 %s cells, one-step-theta, "
                                       "CG abs tol 1e-10 + %s" % ("x".join(map(str, reps)), args.precond),
                           "n_dofs": prob.n_dofs, "nnz_scalar": h.nnz(), "cg_iterations": its,
                           "multigrid_levels": [q.mesh.reps for q in H.problems] if H else None,
                           "multigrid_levels_replicated": H.replicated if H else None,
                           "host_mesh_s": t_mesh, "gf_create_s": t_create,
                           "assemble_once_s": t_asm},
                "roofline": {"bound": "hbm", "kernel": "spmv_tma_kernel<3> (system matrix)",
                             "achieved": nbytes / (avg_ms * 1e-3) / 1e9 if avg_ms else None,
                             "peak": peak, "unit": "GB/s", "peak_source": src,
                             "frac": (nbytes / (avg_ms * 1e-3) / 1e9 / peak) if avg_ms else None,
                             "bytes_per_launch": nbytes, "avg_launch_ms": avg_ms,
                             "launches": int(prof["spmv_launches"]), "standalone_launch_ms": ms_spmv},
                "gpu_launches": int(prof["kernel_launches"]),
                "gpu_mem_gb": torch.cuda.mem_get_info(lr)[1] / 1e9 - torch.cuda.mem_get_info(lr)[0] / 1e9}
        if comm:
            line["comm"] = {"transport": comm.transport()[0]}
        print(json.dumps(line), flush=True)
    (H or h).close()
    if comm:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
