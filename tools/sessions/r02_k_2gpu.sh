mkdir -p gpurun_out
cd /root/repo
echo "== pytest multigrid (+ semi-coarsening), multirank"
timeout 600 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_multirank.py -m gpu -q --durations=4 > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
tail -12 gpurun_out/r02k_pytest.log
echo "== bench --gpus 2 (weak only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 6 --warmup 3 --no-strong > gpurun_out/r02k_bench_n2.json 2> gpurun_out/r02k_bench_n2.err; tail -3 gpurun_out/r02k_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench_n2.json"))
print("N=2 value %.2f M, cg its %d solves %d, ms/cg-it %.2f" % (d["value"]/1e6, d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"], d["config"]["ms_per_cg_iteration"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print(d["config"]["multigrid_levels"], d["config"]["multigrid_levels_replicated"])
PY
