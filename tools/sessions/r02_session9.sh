#!/bin/bash
# Round 2, GPU session 9: modulation GEMV split over a side stream, a slice of the MLP-up GEMM in the attention tail of
# FULL-step single blocks, refitted width / kernel choice, attn_poly = 2 by default - fast tests, racecheck of the new
# stream wiring, step times per knob, bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s9_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s9_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s9_tests_fast.log | tail -12
timeout -k 10 420 compute-sanitizer --tool racecheck --error-exitcode 7 --launch-timeout 0 \
  python -m pytest -q -p no:cacheprovider -m gpu tests/test_flux_parity_gpu.py::test_tiny_flux_default_schedule \
  > gpurun_out/s9_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s9_sanitizer_racecheck.log | tail -3
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_NOP=1"
  run_steps "RGE_SPLIT_MOD=0"
  run_steps "RGE_FILL_ATTN_TAIL=0"
  run_steps "RGE_NOP=2"
} > gpurun_out/s9_step_variants.log 2>&1
cat gpurun_out/s9_step_variants.log
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; tail -c 1800 gpurun_out/s9_bench.json
timeout -k 10 300 python tools/attn_bench.py --quick > gpurun_out/s9_attn_bench.log 2>&1; cut -c1-700 gpurun_out/s9_attn_bench.log
