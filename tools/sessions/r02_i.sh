mkdir -p gpurun_out
cd /root/repo
echo "== K1 probe"; timeout 120 python tools/k1_probe.py 2>&1 | tail -1
echo "== pytest subset"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_reference_pins.py tests/test_gpu_baseline_configs.py tests/test_gpu_multigrid.py -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
tail -5 gpurun_out/r02i_pytest.log
