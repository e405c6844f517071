# short session: two-ring SpMV tests (bitwise vs the single-ring and LDG kernels), then the probe
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_zz_spmv_two_ring.py -q -x > gpurun_out/s15_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s15_pytest.log
tail -15 gpurun_out/s15_pytest.log
timeout 100 python tools/f32_probe.py > gpurun_out/s15_probe.json 2> gpurun_out/s15_probe.err; echo "probe exit $?"
cut -c1-2500 gpurun_out/s15_probe.json
tail -3 gpurun_out/s15_probe.err
