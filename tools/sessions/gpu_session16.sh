# last GPU seconds of the round: the default bench line with the new variants
mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/s16_bench_n1.json 2> gpurun_out/s16_bench_n1.err; echo "bench exit $?"
cut -c1-300 gpurun_out/s16_bench_n1.json
tail -3 gpurun_out/s16_bench_n1.err
