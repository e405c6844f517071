# 1-GPU: how do the CG iteration counts grow with the flap length (the weak-scaling geometries on one GPU)?
mkdir -p gpurun_out
for L in 288 576 1152; do
  timeout 900 python bench.py --layers $L --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s5_bench_layers$L.json 2> gpurun_out/s5_bench_layers$L.err
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/s5_bench_layers$L.json"))
    c = b["config"]
    print("layers $L dofs", c["n_dofs"], "solves", c["newton_solves_in_timed_region"], "cg", c["cg_iterations_in_timed_region"], "value %.2fM" % (b["value"]/1e6), b["phase_ms_per_newton_solve"])
except Exception as e:
    print("layers $L failed", e)
PY
  tail -2 gpurun_out/s5_bench_layers$L.err
done
export GF_PROFILE_RUN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nl_cells_kernel|scatter_matrix_kernel" -c 4 \
    -f -o gpurun_out/s5_prof_asm python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants > gpurun_out/s5_prof_asm.log 2>&1
ls -la gpurun_out | tail -8
