# short session (few GPU-minutes left): the new FP32-V-cycle and output-path tests, then the probe
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_zz_output.py tests/test_gpu_zz_mg_f32.py -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s14_pytest.log
tail -5 gpurun_out/s14_pytest.log
timeout 120 python tools/f32_probe.py > gpurun_out/s14_f32_probe.json 2> gpurun_out/s14_f32_probe.err; echo "probe exit $?"
cut -c1-1500 gpurun_out/s14_f32_probe.json
tail -3 gpurun_out/s14_f32_probe.err
