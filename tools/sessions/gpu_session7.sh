# N-GPU session (N = $NG): parity of the partitioned path, then the bench
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
export GF_COMM_VERBOSE=1
timeout 600 $TR tools/mgpu_check.py > gpurun_out/s7_mgpu_check_n$NG.log 2>&1; echo "exit $?" >> gpurun_out/s7_mgpu_check_n$NG.log
grep -h "mgpu_check\|transport\|exit\|Error\|error" gpurun_out/s7_mgpu_check_n$NG.log | head -30
timeout 600 $TR bench.py --gpus $NG --steps 4 --warmup 3 > gpurun_out/s7_bench_n$NG.json 2> gpurun_out/s7_bench_n$NG.err
cat gpurun_out/s7_bench_n$NG.json; tail -3 gpurun_out/s7_bench_n$NG.err
if [ "$NCCL_TOO" = "1" ]; then
GF_COMM_P2P=0 timeout 600 $TR bench.py --gpus $NG --steps 4 --warmup 3 > gpurun_out/s7_bench_n${NG}_nccl.json 2> gpurun_out/s7_bench_n${NG}_nccl.err
cat gpurun_out/s7_bench_n${NG}_nccl.json
fi
