"""TEST INFRASTRUCTURE ONLY (tests/test_emulated_library.py): two emulated ranks, one of which makes
a collective library call ALONE. The peer-window wait must end with GF_ERR_NCCL after
GF_P2P_TIMEOUT_S instead of hanging, and the error must be sticky (later calls fail at once)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch.distributed as dist
    from dealii_adapter_b200 import build, capi
    emu = os.environ["GF_TEST_EMU_LIB"]
    assert os.path.basename(emu) == "libgraftfem_emu.so"
    build.LIB_CUDA, capi._lib = emu, None
    capi.lib()
    import mgpu_worker as w
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")

    def all_gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = capi.Comm.from_ipc(rank, world, 0, all_gather, share_device=False)
    prob, model, precond, _ = w.make_case("nl_jacobi")
    h = capi.Handle(prob, device=0, partition=prob.mesh.partition(1, world, rank), comm=comm,
                    slab_axis=1)
    h.set_traction(np.tile([1500.0, 0.0, 100.0], h.n_iface_nodes))
    h.nl_begin_step()
    r0 = h.nl_newton_assemble()                     # every rank: fine
    assert r0 > 0
    report = {"rank": rank}
    if rank == 0:
        t0 = time.time()
        try:
            h.nl_newton_assemble()                  # rank 0 ALONE: its peers never arrive
            report["first"] = "returned"
        except capi.GraftError as e:
            report["first"] = (e.code, str(e), time.time() - t0)
        t0 = time.time()
        try:
            h.nl_newton_assemble()
            report["second"] = "returned"
        except capi.GraftError as e:
            report["second"] = (e.code, str(e), time.time() - t0)
    reports = [None] * world
    dist.all_gather_object(reports, report)         # rank 1 waits here, its window stays mapped
    if rank == 0:
        import pickle
        with open(sys.argv[1], "wb") as f:
            pickle.dump(reports[0], f)
    os._exit(0)                                     # the communicator is in its error state


if __name__ == "__main__":
    main()
