#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
export GF_PROFILE_RUN=1
TAG=${TAG:-r01}
# 1) launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-14000} --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants \
    > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# 2) full capture of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:"spmv_tma_kernel|spmv_kernel" -s ${SPMV_SKIP:-60} -c 6 \
    -f -o gpurun_out/${TAG}_prof_spmv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants \
    > gpurun_out/${TAG}_prof_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nl_cells_kernel|scatter_matrix_kernel|mf_apply|mf_setup|cg_update_kernel" -c 6 \
    -f -o gpurun_out/${TAG}_prof_asm python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-variants \
    > gpurun_out/${TAG}_prof_asm.log 2>&1
ls -la gpurun_out
