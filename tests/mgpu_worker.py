"""Worker of tests/test_gpu_multirank.py (one process per rank, started by torch.distributed.run).

Runs coupled windows of both solver classes on a slab-partitioned mesh and writes what rank 0
gathered (Newton table incl. CG iteration counts and residuals, interface displacement of every
written step, transport counters) to --out. Two bootstraps:
  --mode ipc   torch.distributed "gloo" + gf_comm_ipc_begin/_finish: the library's peer-window
               transport with the 64-byte window handles moved by the host. Ranks share GPUs when
               there are fewer devices than ranks (device = local_rank % device_count) - that is
               how the suite covers P = 2, 4 on the one-GPU test box.
  --mode nccl  one GPU per rank, NCCL communicator (gf_comm_create), as bench.py does.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (model, preconditioner, reps, replicate_below_dofs)
    "nl_jacobi": ("neo-Hookean", "jacobi", [2, 8, 2], None),
    "lin_jacobi": ("linear", "jacobi", [2, 8, 2], None),
    "nl_mg": ("neo-Hookean", "mg", [4, 32, 4], 80000),          # small levels replicated (default)
    "lin_mg": ("linear", "mg", [4, 32, 4], 80000),
    "nl_mg_partitioned_coarse": ("neo-Hookean", "mg", [4, 32, 4], 0),   # halos on every level
}
# run only on request (--cases): first hardware run in tests/test_zz_gpu_high_degree.py; the *_small
# cases are sized for the CPU emulation of the library (tests/test_emulated_library.py)
EXTRA_CASES = {
    # name: (model, preconditioner, reps, replicate_below_dofs, polynomial degree)
    "nl_jacobi_q3": ("neo-Hookean", "jacobi", [1, 6, 1], None, 3),
    "lin_jacobi_q3": ("linear", "jacobi", [1, 6, 1], None, 3),
    "nl_mg_small": ("neo-Hookean", "mg", [2, 8, 2], 80000, 2),             # coarse level replicated
    "lin_mg_small": ("linear", "mg", [2, 8, 2], 80000, 2),
    "nl_mg_small_partitioned_coarse": ("neo-Hookean", "mg", [2, 8, 2], 0, 2),   # halos on both levels
    "lin_mg_small_partitioned_coarse": ("linear", "mg", [2, 8, 2], 0, 2),
}
N_STEPS = 2
LOAD = (1500.0, 0.0, 100.0)


def make_case(name):
    from dealii_adapter_b200.problem import SolverParameters, make_problem
    degree = 2
    if name in CASES:
        model, precond, reps, rep_below = CASES[name]
    else:
        model, precond, reps, rep_below, degree = EXTRA_CASES[name]
    p = SolverParameters(model=model, type_lin="CG", poly_degree=degree, scenario="PF", delta_t=0.01,
                         mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-6, max_iterations_lin=2.0)
    return make_problem(p, 3, reps=reps, numbering="lexicographic"), model, precond, rep_below


def run_case(name, world, rank, device, comm):
    """Returns (history, written) of this rank; history rows = the Newton table (nonlinear) or the
    (iterations, residual) pairs of the linear model."""
    from dealii_adapter_b200 import capi, multigrid, solvers
    prob, model, precond, rep_below = make_case(name)
    H = None
    if precond == "mg":
        H = multigrid.Hierarchy(prob, device=device, world=world, rank=rank, comm=comm, axis=1,
                                replicate_below_dofs=rep_below)
        h = H.fine
    else:
        part = prob.mesh.partition(1, world, rank) if world > 1 else None
        h = capi.Handle(prob, device=device, partition=part, comm=comm, slab_axis=1)
    n_if = h.n_iface_nodes
    buf = np.tile(LOAD, n_if)
    fp = solvers.FakeParticipant(3, N_STEPS, prob.params.delta_t, lambda t, it: buf)
    cls = solvers.Solid if model == "neo-Hookean" else solvers.ElastoDynamics
    s = cls(prob, fp, handle=h)
    s.adapter.n_interface_nodes = n_if
    s.adapter.interface_nodes_ids = np.arange(n_if, dtype=np.int32)
    if model == "linear":
        h.lin_assemble_once()
    for k in range(N_STEPS):
        s.step()
    vis = getattr(h, "iface_visible", None)
    written = []
    for (w, it, data) in fp.written:
        full = np.full(prob.n_iface_nodes * 3, np.nan)
        if vis is None:
            full[:] = data
        else:
            full[np.repeat(vis, 3)] = data
        written.append(full)
    hist = [np.array(rows, dtype=np.float64) for rows in s.history] if model == "neo-Hookean" \
        else [np.array(s.history, dtype=np.float64)]
    levels = (H.n_levels, list(H.replicated)) if H else (1, [False])
    (H or h).close()
    return hist, np.array(written), levels


def _watchdog(seconds):
    """A rank that outlives its launcher would keep the GPU busy for everything that runs later:
    every worker ends itself after `seconds`, whatever happens to torchrun."""
    import threading
    import time

    def run():
        time.sleep(seconds)
        sys.stderr.write("mgpu_worker: deadline of %d s exceeded, exiting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=run, daemon=True).start()


def main():
    _watchdog(int(os.environ.get("GF_WORKER_DEADLINE_S", "200")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--cases", default=",".join(CASES))
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    import torch
    import torch.distributed as dist
    from dealii_adapter_b200 import capi
    n_dev = torch.cuda.device_count()
    emu = os.environ.get("GF_TEST_EMU_LIB")
    if emu:
        # tests/test_emulated_library.py: the CPU emulation build of the library, every rank its
        # own "device"; the peer windows are shared-memory files (tests/cuda_emu/cuda_runtime.h)
        from dealii_adapter_b200 import build
        assert args.mode == "ipc" and os.path.basename(emu) == "libgraftfem_emu.so"
        build.LIB_CUDA, capi._lib = emu, None
        capi.lib()
        n_dev = world
    if args.mode == "nccl":
        device = lr
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, device)
    else:
        device = 0 if emu else lr % n_dev
        dist.init_process_group("gloo")

        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        comm = capi.Comm.from_ipc(rank, world, device, all_gather, share_device=n_dev < world)
    result = {"world": world, "mode": args.mode, "shared_device": n_dev < world}
    import time
    for name in args.cases.split(","):
        t0 = time.time()
        if rank == 0:
            print("mgpu_worker: case %s on %d ranks (%s)..." % (name, world, args.mode), flush=True)
        hist, written, levels = run_case(name, world, rank, device, comm)
        if rank == 0:
            print("mgpu_worker: case %s done in %.1f s" % (name, time.time() - t0), flush=True)
        allw = [None] * world
        dist.all_gather_object(allw, written)
        allh = [None] * world
        dist.all_gather_object(allh, hist)
        if rank == 0:
            # every interface node is visible on at least one rank; ranks that see the same node
            # must agree bit for bit (ghost values are copies of the owner's)
            merged = np.full_like(allw[0], np.nan)
            for w in allw:
                m = ~np.isnan(w)
                both = m & ~np.isnan(merged)
                assert np.array_equal(w[both], merged[both]), "ranks disagree on a shared node"
                merged[m] = w[m]
            assert not np.isnan(merged).any()
            for h in allh[1:]:     # the scalars of the Newton table are replicated
                assert all(np.array_equal(a, b) for a, b in zip(h, allh[0])), "histories differ"
            result[name] = {"written": merged, "history": allh[0], "levels": levels}
    if rank == 0:
        result["transport"] = comm.transport()
        import pickle
        with open(args.out, "wb") as f:
            pickle.dump(result, f)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
