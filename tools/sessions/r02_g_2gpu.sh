# Round 2, GPU call G (2 GPUs): K1 with 8 producer warps, parity subset, multi-rank test with the
# NCCL bootstrap on 2 devices, bench.py --gpus 2 (weak cfg3 + strong cfg4).
mkdir -p gpurun_out
cd /root/repo
echo "== K1 probe"; timeout 120 python tools/k1_probe.py 2>&1 | tail -1
echo "== pytest subset"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_reference_pins.py tests/test_gpu_baseline_configs.py tests/test_gpu_multirank.py tests/test_gpu_multigrid.py -m gpu -q --durations=5 > gpurun_out/r02g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_pytest.log
tail -12 gpurun_out/r02g_pytest.log
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; tail -5 gpurun_out/r02g_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02g_bench_n2.json"))
print("N=2 value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d, ms/cg-it %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"], d["config"]["ms_per_cg_iteration"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print(json.dumps(d.get("comm")))
s = d.get("strong_scaling") or {}
print("strong:", {k: s.get(k) for k in ("value", "ms_per_step", "cg_iterations", "ms_per_cg_iteration", "error")}, (s.get("roofline") or {}).get("frac"))
PY
