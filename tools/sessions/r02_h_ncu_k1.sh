mkdir -p gpurun_out
cd /root/repo
GF_PROFILE_RUN=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"nl_cells_kernel" -c 1 -f -o gpurun_out/r02_prof_k1b python tools/ncu_probe.py > gpurun_out/r02_prof_k1b.log 2>&1; tail -2 gpurun_out/r02_prof_k1b.log
