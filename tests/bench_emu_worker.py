"""TEST INFRASTRUCTURE ONLY: one rank of `bench.py --gpus N` on the CPU emulation build of the
library (tests/test_emulated_library.py launches N of these with torch.distributed.run).

bench.py is executed unchanged; what it asks of the machine is redirected around it:
  * capi is bound to the emulation build (GF_TEST_EMU_LIB),
  * torch.distributed starts with the gloo backend instead of nccl, and the few tensors bench.py
    creates on "cuda" (the NCCL id broadcast, the max-over-ranks of the timings) live on the CPU,
  * capi.Comm(unique_id, ...) becomes the library's NCCL-free bootstrap (gf_comm_ipc_*), the
    64-byte window handles travelling through gloo.
Everything else - Hierarchy per rank, the slab partition, every library call of every rank, the
collectives inside them, the order of bench.py's own barriers and all-reduces - is the real thing.
The point: a rank-asymmetric call sequence (what hung one 8-GPU run of round 2) deadlocks HERE,
on the CPU, within the peer-window timeout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from dealii_adapter_b200 import build, capi
    emu = os.environ["GF_TEST_EMU_LIB"]
    assert os.path.basename(emu) == "libgraftfem_emu.so"
    build.LIB_CUDA, capi._lib = emu, None
    capi.lib()
    import bench

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo")
    for name in ("zeros", "tensor"):
        real = getattr(torch, name)
        setattr(torch, name, (lambda real: lambda *a, **kw: real(*a, **{k: v for k, v in kw.items()
                                                                         if k != "device"}))(real))
    real_comm = capi.Comm

    class CommOverGloo(real_comm):
        """capi.Comm(unique_id, rank, n_ranks, device) -> the NCCL-free bootstrap"""
        def __new__(cls, unique_id, rank, n_ranks, device):
            def all_gather(b):
                out = [None] * n_ranks
                dist.all_gather_object(out, b)
                return out
            return real_comm.from_ipc(rank, n_ranks, 0, all_gather, share_device=False)

        def __init__(self, *a):
            pass

        @staticmethod
        def unique_id():
            return bytes(128)
    capi.Comm = CommOverGloo
    bench.CFG4_REPS = tuple(int(x) for x in os.environ.get("GF_TEST_CFG4_REPS", "2,4,2").split(","))
    bench.WEAK_REPS = {}
    bench.ClockSampler.start = lambda self: None
    bench.ClockSampler.stop = lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                                            "samples": 0}
    bench.main()


if __name__ == "__main__":
    main()
