mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s8_pytest.log
tail -4 gpurun_out/s8_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-variants > gpurun_out/s8_bench_n1.json 2> gpurun_out/s8_bench_n1.err
cat gpurun_out/s8_bench_n1.json | cut -c1-400; tail -3 gpurun_out/s8_bench_n1.err
timeout 900 python tools/bench_cfg4.py --steps 3 --warmup 1 > gpurun_out/s8_cfg4_n1.json 2> gpurun_out/s8_cfg4_n1.err
cat gpurun_out/s8_cfg4_n1.json; tail -5 gpurun_out/s8_cfg4_n1.err
