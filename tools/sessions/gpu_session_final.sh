# round-end evidence on one B200: tests, both bench arms, ncu launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > gpurun_out/f_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -3 gpurun_out/f_pytest.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
timeout 600 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
cat gpurun_out/f_bench_n1.json | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
export GF_PROFILE_RUN=1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/f_launches.csv $B > gpurun_out/f_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"spmv_tma_kernel" -s 8 -c 2 -f -o gpurun_out/f_prof_spmv $B > gpurun_out/f_prof_spmv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nl_cells_kernel|scatter_matrix_kernel|coarse_cheb_kernel|cheb_step_kernel" -c 6 -f -o gpurun_out/f_prof_asm $B > gpurun_out/f_prof_asm.log 2>&1
ls -la gpurun_out | grep " f_"
