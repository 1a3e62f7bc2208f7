#!/bin/bash
# Round 2, GPU session 10: K/V split of the ragged last query tile (REGION attention), racecheck after the redundant
# CTA barrier in gemm2, whole GPU suite, step times with / without the split.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s10_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s10_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s10_tests_fast.log | tail -12
timeout -k 10 300 python tools/attn_bench.py --quick > gpurun_out/s10_attn_bench.log 2>&1; cut -c1-900 gpurun_out/s10_attn_bench.log
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_NOP=1"
  run_steps "RGE_ATTN_SPLIT=0"
  run_steps "RGE_NOP=2"
  run_steps "RGE_ATTN_SPLIT=0 RGE_NOP=3"
} > gpurun_out/s10_step_variants.log 2>&1
grep -v SKIP gpurun_out/s10_step_variants.log
for tool in racecheck memcheck; do
  timeout -k 10 420 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 \
    python -m pytest -q -p no:cacheprovider -m gpu tests/test_flux_parity_gpu.py::test_tiny_flux_default_schedule \
    "tests/test_kernels_gpu.py::test_attention_kv_split_of_the_ragged_last_query_tile[1537-1100-24]" \
    > gpurun_out/s10_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s10_sanitizer_$tool.log | tail -3
done
timeout -k 10 1200 python -m pytest -q -rA -p no:cacheprovider -m gpu tests --deselect tests/test_kernels_gpu.py \
  --deselect tests/test_trim_last_gpu.py --deselect tests/test_flux_parity_gpu.py --deselect tests/test_flux_fullsize_gpu.py \
  > gpurun_out/s10_tests_rest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s10_tests_rest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s10_tests_rest.log | tail -12
