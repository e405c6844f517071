# 2-GPU session: parity of the partitioned path with both transports, then the bench with both
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
export GF_COMM_VERBOSE=1
timeout 600 $TR tools/mgpu_check.py > gpurun_out/s3_mgpu_check_p2p.log 2>&1; echo "exit $?" >> gpurun_out/s3_mgpu_check_p2p.log
GF_COMM_P2P=0 timeout 600 $TR tools/mgpu_check.py > gpurun_out/s3_mgpu_check_nccl.log 2>&1; echo "exit $?" >> gpurun_out/s3_mgpu_check_nccl.log
grep -h "mgpu_check\|transport\|exit\|Error\|error" gpurun_out/s3_mgpu_check_p2p.log gpurun_out/s3_mgpu_check_nccl.log | head -40
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/s3_bench_n2_p2p.json 2> gpurun_out/s3_bench_n2_p2p.err
GF_COMM_P2P=0 timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/s3_bench_n2_nccl.json 2> gpurun_out/s3_bench_n2_nccl.err
cat gpurun_out/s3_bench_n2_p2p.json gpurun_out/s3_bench_n2_nccl.json
tail -5 gpurun_out/s3_bench_n2_p2p.err
