mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/s4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4_pytest.log
tail -25 gpurun_out/s4_pytest.log
timeout 600 python tools/spmv_tune.py > gpurun_out/s4_spmv_tune.jsonl 2> gpurun_out/s4_spmv_tune.err
cat gpurun_out/s4_spmv_tune.jsonl; tail -3 gpurun_out/s4_spmv_tune.err
timeout 900 python tools/mg_sweep.py > gpurun_out/s4_mg_sweep.jsonl 2> gpurun_out/s4_mg_sweep.err
tail -3 gpurun_out/s4_mg_sweep.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err
cat gpurun_out/s4_bench_n1.json
