# Round 2, first GPU call (1 GPU, ~6 min): everything staged at the end of round 1 that has not run on
# hardware yet. Outputs in gpurun_out/r02a_*.
mkdir -p gpurun_out
export GF_TEST_EXPERIMENTAL=1
# 1) the experimental paths: fused-dot on 16 consumer warps (kinds 3/4), transposed reduction
#    (kind 6), all-FP32 V-cycle operator (precision 2)
timeout 200 python -m pytest tests/test_gpu_zz_spmv_two_ring.py tests/test_gpu_zz_mg_f32.py -q > gpurun_out/r02a_experimental_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_experimental_tests.log
tail -4 gpurun_out/r02a_experimental_tests.log
# 2) launch times of every kernel kind incl. 6 and the all-FP32 operator on the cfg3 tangent
timeout 60 python tools/spmv_kinds_probe.py > gpurun_out/r02a_kinds.jsonl 2> gpurun_out/r02a_kinds.err; cat gpurun_out/r02a_kinds.jsonl
# 3) the bench line with candidate defaults (short runs)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
for cfg in "--spmv-kernel 0 --mg-precision 0" "--spmv-kernel 3 --mg-precision 0" "--spmv-kernel 6 --mg-precision 0" \
           "--spmv-kernel 6 --mg-precision 1" "--spmv-kernel 6 --mg-precision 2" "--spmv-kernel 3 --mg-precision 2" \
           "--spmv-kernel 0 --mg-refresh 2" "--spmv-kernel 0 --mg-refresh 4"; do
  tag=$(echo $cfg | tr -d ' -' )
  timeout 90 $B $cfg > gpurun_out/r02a_bench_$tag.json 2> gpurun_out/r02a_bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02a_bench_$tag.json"))
    print("$cfg", "%.2f M DoFs/s" % (d["value"] / 1e6), "spmv %.3f ms" % d["roofline"]["avg_launch_ms"],
          "cg its", d["config"]["cg_iterations_in_timed_region"])
except Exception as e:
    print("$cfg", "FAILED", e)
PY
done
