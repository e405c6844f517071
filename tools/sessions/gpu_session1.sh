mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1_pytest.log
timeout 600 python bench.py > gpurun_out/s1_bench_mg.json 2> gpurun_out/s1_bench_mg.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s1_bench_ref.json 2> gpurun_out/s1_bench_ref.err
TAG=s1 NCU_COUNT=6000 timeout 900 bash tools/profile.sh > gpurun_out/s1_profile.log 2>&1
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_bench_mg.json
