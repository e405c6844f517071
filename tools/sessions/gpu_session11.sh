mkdir -p gpurun_out
for R in 24,288,48; do
  timeout 900 python bench.py --reps $R --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s11_bench_reps_$R.json 2> gpurun_out/s11_bench_reps_$R.err
  tail -2 gpurun_out/s11_bench_reps_$R.err
done
timeout 900 python bench.py --reps 48,288,48 --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s11_bench_reps_48,288,48.json 2> gpurun_out/s11_bench_reps_48,288,48.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s11_bench_*.json")):
    try:
        b = json.load(open(f)); c = b["config"]
        print(f, "dofs", c["n_dofs"], "solves", c["newton_solves_in_timed_region"], "cg", c["cg_iterations_in_timed_region"], "value %.2fM" % (b["value"]/1e6), "levels", c["multigrid_levels"])
        print("   ", {k: round(v, 1) for k, v in b["phase_ms_per_newton_solve"].items() if k != "note"})
    except Exception as e:
        print(f, "failed", e)
PY
