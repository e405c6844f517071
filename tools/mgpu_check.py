"""Multi-GPU parity check, launched with torchrun (one rank per GPU): the slab-partitioned path
(halo exchange + all-reduced dots/norms) must reproduce the single-GPU run — same Newton counts,
interface displacement to 1e-9 relative."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dealii_adapter_b200 import capi, multigrid, solvers
from dealii_adapter_b200.problem import SolverParameters, make_problem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm = capi.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, lr)
# "mg": all coarse levels slab-partitioned (halo exchanges on every level);
# "mg-replicated": the small levels replicated on every rank (the default of multigrid.Hierarchy)
for model, type_lin, precond in (("neo-Hookean", "CG", "jacobi"), ("linear", "CG", "jacobi"),
                                 ("neo-Hookean", "CG", "mg"), ("linear", "CG", "mg"),
                                 ("neo-Hookean", "CG", "mg-replicated"),
                                 ("linear", "CG", "mg-replicated")):
    p = SolverParameters(model=model, type_lin=type_lin, poly_degree=2, scenario="PF", delta_t=0.01,
                         mu=0.5e6, nu=0.4, rho=1000.0, tol_lin=1e-8, max_iterations_lin=2.0)
    reps = [3, 4 * world, 2] if precond == "jacobi" else [4, 8 * world, 4]
    prob = make_problem(p, 3, reps=reps, numbering="lexicographic")
    n = prob.n_iface_nodes
    load = np.array([1500.0, 0.0, 100.0])
    H = None
    if precond.startswith("mg"):
        H = multigrid.Hierarchy(prob, device=lr, world=world, rank=rank, comm=comm, axis=1,
                                replicate_below_dofs=0 if precond == "mg" else 80000)
        h = H.fine
    else:
        part = prob.mesh.partition(1, world, rank)
        h = capi.Handle(prob, device=lr, partition=part, comm=comm)
    buf = np.tile(load, h.n_iface_nodes)
    fp = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: buf)
    cls = solvers.Solid if model == "neo-Hookean" else solvers.ElastoDynamics
    s = cls(prob, fp, handle=h)
    s.adapter.n_interface_nodes = h.n_iface_nodes
    s.adapter.interface_nodes_ids = np.arange(h.n_iface_nodes, dtype=np.int32)
    if model == "linear":
        h.lin_assemble_once()
    for k in range(3):
        s.step()
    mine = np.full(n * 3, np.nan)
    vis = np.repeat(h.iface_visible, 3)
    mine[vis] = fp.written[-1][2]
    allv = [None] * world
    dist.all_gather_object(allv, mine)
    if rank == 0:
        Hs = multigrid.Hierarchy(prob, device=lr) if precond.startswith("mg") else None
        hs = Hs.fine if Hs else capi.Handle(prob, device=lr)
        bufs = np.tile(load, n)
        fps = solvers.FakeParticipant(3, 3, p.delta_t, lambda t, it: bufs)
        ss = cls(prob, fps, handle=hs)
        ss.adapter.initialize(prob)
        if model == "linear":
            hs.lin_assemble_once()
        for k in range(3):
            ss.step()
        ref = fps.written[-1][2]
        for r in range(world):
            m = ~np.isnan(allv[r])
            err = np.abs(allv[r][m] - ref[m]).max() / np.abs(ref).max()
            assert err < (1e-9 if precond == "jacobi" else 1e-7), (model, precond, r, err)
        if model == "neo-Hookean":
            assert [len(x) for x in s.history] == [len(x) for x in ss.history], (s.history, ss.history)
        print("mgpu_check %s %s world=%d levels=%d replicated=%s OK (history %s | single-GPU %s)" % (
              model, precond, world, H.n_levels if H else 1, H.replicated if H else None,
              [[r[0] for r in x] for x in s.history] if model == "neo-Hookean" else s.history,
              [[r[0] for r in x] for x in ss.history] if model == "neo-Hookean" else ss.history))
        (Hs or hs).close()
    (H or h).close()
if rank == 0:
    print("mgpu_check transport:", comm.transport())
comm.close()
dist.destroy_process_group()
