# Round 2, GPU call C (1 GPU): ranks sharing device 0 through the NCCL-free bootstrap with stream
# memory-op waits; timing of the cases, then the multirank pytest, K1 symmetric check, short bench.
mkdir -p gpurun_out
cd /root/repo
run_worker () {  # $1 = ranks, $2 = cases
  GF_WORKER_DEADLINE_S=120 timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 \
      --master-port 29531 tests/mgpu_worker.py --out gpurun_out/r02c_p$1.pkl --mode ipc --cases $2 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^W1\|^$" | tail -12
  sleep 1
}
export GF_P2P_TIMEOUT_S=20
echo "== P=2 nl_jacobi"; run_worker 2 nl_jacobi
echo "== P=2 all"; run_worker 2 nl_jacobi,lin_jacobi,nl_mg,lin_mg,nl_mg_partitioned_coarse
echo "== P=4 all"; run_worker 4 nl_jacobi,lin_jacobi,nl_mg,lin_mg,nl_mg_partitioned_coarse
nvidia-smi --query-compute-apps=pid,used_memory --format=csv
echo "== pytest (parity, multirank, K1 symmetric)"
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py tests/test_gpu_zz_reference_pins.py tests/test_gpu_baseline_configs.py tests/test_gpu_zz_spmv_two_ring.py -m gpu -x -q --durations=8 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
nvidia-smi --query-compute-apps=pid,used_memory --format=csv
echo "== bench"
timeout 300 python bench.py --steps 4 --warmup 3 --no-strong > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_bench.json"))
print("value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print(json.dumps(d.get("variants"))[:1200])
PY
