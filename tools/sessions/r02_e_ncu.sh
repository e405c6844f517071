# Round 2, GPU call E (1 GPU): ncu evidence of the shipped defaults.
mkdir -p gpurun_out
cd /root/repo
export GF_PROFILE_RUN=1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"spmv_tma2_kernel|nl_cells_kernel|scatter_matrix_kernel" -c 6 -f -o gpurun_out/r02_prof \
    python tools/ncu_probe.py > gpurun_out/r02_prof.log 2>&1; tail -3 gpurun_out/r02_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants --no-strong \
    > gpurun_out/r02_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/r02_bench_under_ncu.log
ls -la gpurun_out | tail -5
