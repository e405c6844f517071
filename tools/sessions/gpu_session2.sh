mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/s2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s2_pytest.log
tail -15 gpurun_out/s2_pytest.log
timeout 600 python bench.py > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
cat gpurun_out/s2_bench_n1.json
timeout 900 python tools/mg_sweep.py > gpurun_out/s2_mg_sweep.jsonl 2> gpurun_out/s2_mg_sweep.err
export GF_PROFILE_RUN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_tma_kernel" -s 8 -c 3 \
    -f -o gpurun_out/s2_prof_spmv_fine python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants \
    > gpurun_out/s2_prof_spmv_fine.log 2>&1
ls -la gpurun_out
