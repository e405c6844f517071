# Round 2, GPU call D (1 GPU): multirank cases with visible timing, pytest subset, K1 probe, bench.
mkdir -p gpurun_out
cd /root/repo
export GF_P2P_TIMEOUT_S=20
run_worker () {
  GF_WORKER_DEADLINE_S=150 timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 \
      --master-port 2953$1 tests/mgpu_worker.py --out gpurun_out/r02d_p$1.pkl --mode ipc --cases $2 2>&1 | grep "mgpu_worker\|Error\|error\|assert" | head -30
  sleep 1
}
echo "== P=2 all"; run_worker 2 nl_jacobi,lin_jacobi,nl_mg,lin_mg,nl_mg_partitioned_coarse
echo "== P=4 all"; run_worker 4 nl_jacobi,lin_jacobi,nl_mg,lin_mg,nl_mg_partitioned_coarse
echo "== K1 probe"; timeout 120 python tools/k1_probe.py 2>&1 | tail -2
echo "== pytest"
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py tests/test_gpu_zz_reference_pins.py tests/test_gpu_baseline_configs.py tests/test_gpu_zz_spmv_two_ring.py tests/test_gpu_multigrid.py tests/test_gpu_matfree.py tests/test_gpu_zz_mg_f32.py tests/test_gpu_zz_output.py tests/test_gpu_flap_mid.py -m gpu -q --durations=8 > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
tail -40 gpurun_out/r02d_pytest.log
echo "== bench"
timeout 300 python bench.py --steps 4 --warmup 3 --no-strong > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -3 gpurun_out/r02d_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_bench.json"))
print("value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print(json.dumps(d.get("variants"))[:600])
print(json.dumps(d.get("cpu_baseline"))[:900])
PY
