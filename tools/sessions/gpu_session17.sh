# last GPU seconds: ncu --set full of the four SpMV variants (single/two ring x FP64/FP32 copy)
mkdir -p gpurun_out
timeout 75 ncu --set full --clock-control none --profile-from-start off --import-source on -f -o gpurun_out/s17_spmv_variants python tools/spmv_ncu_probe.py > gpurun_out/s17_ncu.log 2>&1; echo "ncu exit $?"
tail -4 gpurun_out/s17_ncu.log
ls -la gpurun_out | grep s17
