mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s9_pytest.log
tail -6 gpurun_out/s9_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s9_bench_n1.json 2> gpurun_out/s9_bench_n1.err
cat gpurun_out/s9_bench_n1.json | cut -c1-300; tail -3 gpurun_out/s9_bench_n1.err
GF_COARSE_SOLVER=0 timeout 600 python bench.py --no-cpu-baseline --no-variants > gpurun_out/s9_bench_n1_nocs.json 2> gpurun_out/s9_bench_n1_nocs.err
for R in 24,144,48 48,144,48 48,288,48; do
  timeout 900 python bench.py --reps $R --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s9_bench_reps_$R.json 2> gpurun_out/s9_bench_reps_$R.err
  tail -2 gpurun_out/s9_bench_reps_$R.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s9_bench_*.json")):
    try:
        b = json.load(open(f)); c = b["config"]
        print(f, "dofs", c["n_dofs"], "solves", c["newton_solves_in_timed_region"], "cg", c["cg_iterations_in_timed_region"], "value %.2fM" % (b["value"]/1e6), "levels", c["multigrid_levels"])
        print("   ", {k: round(v, 1) for k, v in b["phase_ms_per_newton_solve"].items() if k != "note"})
    except Exception as e:
        print(f, "failed", e)
PY
