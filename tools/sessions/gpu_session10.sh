# 8-GPU session: the driver's scaling command at N=8, then cfg4 (51M DoFs) strong-scaled over 8 GPUs
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
export GF_COMM_VERBOSE=1
free -g | head -2; nproc
timeout 600 $TR bench.py --gpus $NG --steps 4 --warmup 3 > gpurun_out/s10_bench_n$NG.json 2> gpurun_out/s10_bench_n$NG.err
cat gpurun_out/s10_bench_n$NG.json; tail -3 gpurun_out/s10_bench_n$NG.err
FREE=$(free -g | awk '/Mem:/ {print $7}')
if [ "$FREE" -gt 250 ]; then
  timeout 600 $TR tools/bench_cfg4.py --steps 3 --warmup 1 > gpurun_out/s10_cfg4_n$NG.json 2> gpurun_out/s10_cfg4_n$NG.err
  grep '^{' gpurun_out/s10_cfg4_n$NG.json; tail -3 gpurun_out/s10_cfg4_n$NG.err
else
  echo "skipping cfg4: only $FREE GB of host memory available"
fi
