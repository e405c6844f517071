# Round 2, GPU call J (8 GPUs): bench.py --gpus 8 (weak cfg3 + strong cfg4), cfg5 (21.7 M DoFs, 10
# sub-iterations), multi-rank parity test with NCCL-bootstrapped peer windows on real devices.
mkdir -p gpurun_out
cd /root/repo
echo "== bench --gpus 8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02j_bench_n8.json 2> gpurun_out/r02j_bench_n8.err; tail -3 gpurun_out/r02j_bench_n8.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02j_bench_n8.json"))
    print("N=8 value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d, ms/cg-it %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"], d["config"]["ms_per_cg_iteration"]))
    print(json.dumps(d["phase_ms_per_newton_solve"]))
    print(json.dumps(d.get("comm")))
    print(d["config"]["multigrid_levels"], d["config"]["multigrid_levels_replicated"])
    s = d.get("strong_scaling") or {}
    print("strong:", {k: s.get(k) for k in ("value", "ms_per_step", "cg_iterations", "ms_per_cg_iteration", "error", "multigrid_levels_replicated")}, (s.get("roofline") or {}).get("frac"))
except Exception as e:
    print("bench n8 failed", e)
PY
echo "== cfg5"
STEPS=10 WARMUP=10 timeout 600 bash tools/bench_cfg5.sh > gpurun_out/r02j_cfg5_n8.json 2> gpurun_out/r02j_cfg5_n8.err; tail -3 gpurun_out/r02j_cfg5_n8.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02j_cfg5_n8.json"))
    print("cfg5 N=8 value %.2f M e2e %.2f M, n_dofs %d, spmv frac %.3f, cg its %d solves %d, ms/step %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["config"]["n_dofs"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"], d["ms_per_step"]))
    print(json.dumps(d["phase_ms_per_newton_solve"]))
    print(d["config"]["multigrid_levels"], d["config"]["multigrid_levels_replicated"])
except Exception as e:
    print("cfg5 failed", e)
PY
echo "== multirank pytest (nccl on real devices too)"
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; tail -4 gpurun_out/r02j_pytest.log
