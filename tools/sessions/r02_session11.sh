#!/bin/bash
# Round 2, GPU session 11: K/V split generalised to the trailing work units of any attention grid (FULL steps: 5.51
# waves) - kernel tests, isolated attention, step times with / without, racecheck with and without the CTA-pair GEMM.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s11_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s11_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s11_tests_fast.log | tail -12
timeout -k 10 300 python tools/attn_bench.py --quick > gpurun_out/s11_attn_bench.log 2>&1; cut -c1-1000 gpurun_out/s11_attn_bench.log
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_NOP=1"
  run_steps "RGE_ATTN_SPLIT=0"
  run_steps "RGE_NOP=2"
  run_steps "RGE_ATTN_SPLIT=0 RGE_NOP=3"
} > gpurun_out/s11_step_variants.log 2>&1
grep -v SKIP gpurun_out/s11_step_variants.log
for knobs in "RGE_2CTA_MIN_M=-1" "RGE_2CTA_MIN_M=2048"; do
  env $knobs timeout -k 10 420 compute-sanitizer --tool racecheck --error-exitcode 7 --launch-timeout 0 \
    python -m pytest -q -p no:cacheprovider -m gpu tests/test_flux_parity_gpu.py::test_tiny_flux_default_schedule \
    "tests/test_kernels_gpu.py::test_attention_kv_split_of_the_trailing_work_units[1537-1100-24]" \
    > "gpurun_out/s11_sanitizer_racecheck_${knobs}.log" 2>&1
  echo "racecheck $knobs rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" "gpurun_out/s11_sanitizer_racecheck_${knobs}.log" | tail -2
done
