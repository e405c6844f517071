"""Host-side logic of the partitioned (N>1) path on CPU: slab partition + neighbour exchange lists,
exercised with a world_size-2/3 gloo group (the device path does the same exchange with NCCL
send/recv on packed buffers, csrc/comm.cu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import nl_params
from dealii_adapter_b200.problem import make_problem


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, reps, numbering, out, degree=2):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = make_problem(nl_params(poly_degree=degree), 3, reps=reps, numbering=numbering)
        part = prob.mesh.partition(1, world, rank)
        rng = np.random.RandomState(42)
        g = rng.uniform(-1, 1, prob.n_dofs)              # same global vector on every rank
        loc = g[part.local_to_global].copy()
        loc[part.n_owned_dofs:] = np.nan                  # ghosts unknown before the exchange
        reqs, recv_bufs = [], []
        for k, nbr in enumerate(part.nbr_rank):
            s = torch.from_numpy(loc[part.send_dofs[part.send_ptr[k]:part.send_ptr[k + 1]]].copy())
            r = torch.empty(int(part.recv_ptr[k + 1] - part.recv_ptr[k]), dtype=torch.float64)
            reqs.append(dist.isend(s, int(nbr)))
            reqs.append(dist.irecv(r, int(nbr)))
            recv_bufs.append((k, r))
        for q in reqs:
            q.wait()
        for k, r in recv_bufs:
            loc[part.recv_dofs[part.recv_ptr[k]:part.recv_ptr[k + 1]]] = r.numpy()
        ok = np.array_equal(loc, g[part.local_to_global])
        # all-reduce of an owned-only dot product reproduces the global one
        d = torch.tensor([float(np.dot(loc[:part.n_owned_dofs], loc[:part.n_owned_dofs]))],
                         dtype=torch.float64)
        dist.all_reduce(d)
        ok = ok and abs(d.item() - float(np.dot(g, g))) < 1e-9 * float(np.dot(g, g))
        # owned sets partition the global dofs
        owned = torch.zeros(prob.n_dofs, dtype=torch.int32)
        owned[torch.from_numpy(part.local_to_global[:part.n_owned_dofs].astype(np.int64))] = 1
        dist.all_reduce(owned)
        ok = ok and bool((owned == 1).all())
        # every local cell only references local dofs; owned rows are complete (all cells around an
        # owned node are local)
        cells_g = set(part.local_cell_global.tolist())
        cd = prob.mesh.cell_dofs.reshape(-1, prob.mesh.dofs_per_cell)
        # the local cell -> dof table is the global one in local numbering, entry by entry (the
        # FESystem local order must survive the renumbering, also above degree 2)
        lcd = part.cell_dofs.reshape(-1, prob.mesh.dofs_per_cell)
        ok = ok and np.array_equal(part.local_to_global[lcd], cd[part.local_cell_global])
        owned_g = set(part.local_to_global[:part.n_owned_dofs].tolist())
        for c in range(prob.mesh.n_cells):
            if c not in cells_g and owned_g.intersection(cd[c].tolist()):
                ok = False
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,reps,numbering,degree", [(2, [2, 4, 2], "cellwise", 2),
                                                         (3, [1, 5, 2], "component_wise", 2),
                                                         (2, [1, 4, 2], "cellwise", 3)])
def test_halo_exchange_lists_with_gloo(native_libs, world, reps, numbering, degree):
    out = mp.get_context("spawn").Array("i", [0] * world)
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker,
                                             args=(r, world, port, reps, numbering, out, degree))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert list(out) == [1] * world
