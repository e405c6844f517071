# Round 2, GPU call B (1 GPU): the whole GPU suite after the partition-independent reductions /
# numbering change, the new bench-size parity tests, the multi-rank tests (ranks share device 0),
# then one bench line with variants.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -40 gpurun_out/r02b_pytest.log
timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 3000 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
