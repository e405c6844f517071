mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/s6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s6_pytest.log
tail -12 gpurun_out/s6_pytest.log
timeout 600 python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err
cat gpurun_out/s6_bench_n1.json; tail -3 gpurun_out/s6_bench_n1.err
export GF_PROFILE_RUN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nl_cells_kernel" -c 1 \
    -f -o gpurun_out/s6_prof_cells python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants > gpurun_out/s6_prof_cells.log 2>&1
