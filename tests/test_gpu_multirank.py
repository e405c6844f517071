"""Multi-rank parity inside `pytest -m gpu`: the slab-partitioned path (ghost-DoF halo, reductions,
replicated / partitioned multigrid levels) at P = 2 and 4 ranks against the single-rank run of the
same mesh. The reference is a serial program (adapter.h:152-154), so P = 1 is the truth and every
P must reproduce it:
  * identical Newton iteration counts AND identical CG iteration counts per Newton step,
  * interface displacement of every written step BITWISE equal - all cross-rank sums run over a
    partition-independent tree (csrc/reduce.cuh, gf_desc.slab_axis / dof_global); only the case
    with slab-partitioned coarsest level uses another coarse solver path (launch-by-launch instead
    of the single cooperative launch) and is held to 1e-9.
The ranks are real processes (torch.distributed.run). On a box with fewer GPUs than ranks they
share device 0 through the library's NCCL-free bootstrap (gf_comm_ipc_*; cudaIpc windows work
between processes on one device, NCCL does not), time-sliced by the driver; with enough GPUs the
same worker also runs over NCCL-bootstrapped peer windows (tests/mgpu_worker.py --mode nccl)."""
import os
import pickle
import signal
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(world, mode, out, cases=None, timeout=240):
    env = dict(os.environ, GF_P2P_TIMEOUT_S=os.environ.get("GF_TEST_P2P_TIMEOUT_S", "30"),
               OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "mgpu_worker.py"), "--out", out, "--mode", mode]
    if cases:
        cmd += ["--cases", cases]
    # own process group: a timeout must take the rank processes down too (a killed torchrun
    # leaves its workers behind, and they would keep the GPU busy for every later test)
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    try:
        stdout, _ = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        stdout, _ = proc.communicate()
        raise AssertionError("multi-rank worker timed out after %d s\n%s" % (timeout, stdout[-4000:]))
    assert proc.returncode == 0, stdout[-4000:]
    with open(out, "rb") as f:
        return pickle.load(f)


@pytest.fixture(scope="module")
def single_rank(native_libs):
    """P = 1 through the same worker code in this process."""
    native_libs.build_cuda()
    import mgpu_worker as w
    out = {}
    for name in w.CASES:
        hist, written, levels = w.run_case(name, 1, 0, 0, None)
        out[name] = {"written": written, "history": hist, "levels": levels}
    return out


def _compare(ref, got, name, bitwise):
    a, b = ref[name], got[name]
    assert len(a["history"]) == len(b["history"])
    for ha, hb in zip(a["history"], b["history"]):
        assert ha.shape == hb.shape, (name, "Newton counts differ", ha.shape, hb.shape)
        assert np.array_equal(ha[:, 0], hb[:, 0]), (name, "CG iteration counts", ha[:, 0], hb[:, 0])
        if bitwise:
            assert np.array_equal(ha, hb), (name, "Newton table differs in the last bits")
    if bitwise:
        assert np.array_equal(a["written"], b["written"]), name
    else:
        assert np.abs(a["written"] - b["written"]).max() <= 1e-9 * np.abs(a["written"]).max()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_runs_reproduce_the_single_rank_run(single_rank, tmp_path, world):
    import torch
    import mgpu_worker as w
    got = _spawn(world, "ipc", str(tmp_path / ("ipc%d.pkl" % world)))
    assert got["transport"][0] == "peer_windows" and got["transport"][1] > 0
    for name in w.CASES:
        bitwise = name != "nl_mg_partitioned_coarse"
        if name == "nl_mg_partitioned_coarse":
            # same Newton counts; CG counts may move by one where the coarse solver path differs
            a, b = single_rank[name], got[name]
            assert [h.shape for h in a["history"]] == [h.shape for h in b["history"]]
            assert all(np.abs(x[:, 0] - y[:, 0]).max() <= 1 for x, y in zip(a["history"], b["history"]))
            assert np.abs(a["written"] - b["written"]).max() <= 1e-9 * np.abs(a["written"]).max()
            assert not any(b["levels"][1])
        else:
            _compare(single_rank, got, name, bitwise)
    assert got["nl_mg"]["levels"][0] >= 3 and any(got["nl_mg"]["levels"][1])
    if torch.cuda.device_count() >= world:      # a real multi-GPU box: NCCL bootstrap as well
        got2 = _spawn(world, "nccl", str(tmp_path / ("nccl%d.pkl" % world)))
        for name in ("nl_jacobi", "lin_jacobi", "nl_mg", "lin_mg"):
            _compare(single_rank, got2, name, True)
