mkdir -p gpurun_out
cd /root/repo
echo "== pytest (all gpu tests)"
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest.log
tail -12 gpurun_out/r02l_pytest.log
echo "== bench"
timeout 600 python bench.py --steps 6 --warmup 3 --no-strong > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -3 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_bench.json"))
print("value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print({k: (round(v["value"]/1e6, 2), v.get("ms_per_newton_solve")) for k, v in d.get("variants", {}).items() if isinstance(v, dict) and "value" in v}, d.get("variants", {}).get("error"))
PY
