#!/bin/bash
# Round 2, GPU session 7: attention issue-order variants (early upper half of Q K^T, P in 3 pieces) and the CTA-pair
# GEMM's per-shape tile width (pick_bn2) - kernel tests, isolated attention / GEMM throughput, step times per variant.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s7_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s7_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s7_tests_fast.log | tail -12
timeout -k 10 400 python tools/attn_bench.py > gpurun_out/s7_attn_bench.log 2>&1; cat gpurun_out/s7_attn_bench.log | cut -c1-1400
timeout -k 10 600 python tools/gemm_bench.py --quick > gpurun_out/s7_gemm_bench.log 2>&1; cat gpurun_out/s7_gemm_bench.log | cut -c1-700
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_ATTN_VARIANT=0 RGE_GEMM2_BN=256"
  run_steps "RGE_ATTN_VARIANT=0"
  run_steps "RGE_ATTN_VARIANT=1"
  run_steps "RGE_ATTN_VARIANT=2"
  run_steps "RGE_ATTN_VARIANT=2 RGE_ATTN_POLY=2"
  run_steps "RGE_ATTN_VARIANT=1 RGE_ATTN_POLY=2"
} > gpurun_out/s7_step_variants.log 2>&1
cat gpurun_out/s7_step_variants.log
