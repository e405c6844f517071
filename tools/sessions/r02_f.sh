# Round 2, GPU call F (1 GPU): K1 with 7 consumer warps, Q1 SpMV kinds + ncu, full GPU suite,
# bench line incl. cfg4 strong scaling at N = 1.
mkdir -p gpurun_out
cd /root/repo
echo "== K1 probe"; timeout 120 python tools/k1_probe.py 2>&1 | tail -1
echo "== Q1 kinds"; timeout 200 python tools/q1_spmv_probe.py 2>&1 | tail -5
GF_PROFILE_RUN=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"spmv_tma2_kernel" -c 1 -f -o gpurun_out/r02_prof_q1 python tools/q1_spmv_probe.py --ncu > gpurun_out/r02_prof_q1.log 2>&1; tail -2 gpurun_out/r02_prof_q1.log
echo "== pytest"
timeout 1200 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -14 gpurun_out/r02f_pytest.log
echo "== bench"
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02f_bench.json"))
print("value %.2f M e2e %.2f M, spmv %.3f ms frac %.3f, cg its %d solves %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["config"]["cg_iterations_in_timed_region"], d["config"]["newton_solves_in_timed_region"]))
print(json.dumps(d["phase_ms_per_newton_solve"]))
print(json.dumps(d.get("strong_scaling"))[:1800])
print({k: round(v["value"]/1e6, 2) for k, v in d.get("variants", {}).items() if isinstance(v, dict) and "value" in v})
PY
