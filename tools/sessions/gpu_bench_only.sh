NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $NG --steps 4 --warmup 3 > gpurun_out/s12_bench_n$NG.json 2> gpurun_out/s12_bench_n$NG.err
cat gpurun_out/s12_bench_n$NG.json | cut -c1-1200; tail -3 gpurun_out/s12_bench_n$NG.err
