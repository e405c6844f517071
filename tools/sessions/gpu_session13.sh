mkdir -p gpurun_out
timeout 900 python bench.py --reps 24,288,96 --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s13_bench_reps_24,288,96.json 2> gpurun_out/s13.err
tail -2 gpurun_out/s13.err
python - <<'PY'
import json
b = json.load(open("gpurun_out/s13_bench_reps_24,288,96.json")); c = b["config"]
print("dofs", c["n_dofs"], "solves", c["newton_solves_in_timed_region"], "cg", c["cg_iterations_in_timed_region"], "value %.2fM" % (b["value"]/1e6), "levels", c["multigrid_levels"], b["roofline"]["achieved"])
print("   ", {k: round(v, 1) for k, v in b["phase_ms_per_newton_solve"].items() if k != "note"})
PY
