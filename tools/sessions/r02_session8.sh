#!/bin/bash
# Round 2, GPU session 8: single-pass NORM_ROPE epilogue with early TMEM release, CTA-pair tile-width sweep, per-shape
# kernel choice for REGION-sized GEMMs, exponential offload in-step; ncu source profile of the NORM_ROPE GEMM.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s8_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s8_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s8_tests_fast.log | tail -12
timeout -k 10 400 python tools/gemm2_width_sweep.py > gpurun_out/s8_width_sweep.log 2>&1; cat gpurun_out/s8_width_sweep.log | cut -c1-400
timeout -k 10 600 python tools/gemm_bench.py --quick > gpurun_out/s8_gemm_bench.log 2>&1; head -16 gpurun_out/s8_gemm_bench.log | cut -c1-800
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_NOP=1"
  run_steps "RGE_ATTN_POLY=2"
  run_steps "RGE_ATTN_POLY=3"
  run_steps "RGE_2CTA_MIN_M=2048"
  run_steps "RGE_GEMM2_BN=256"
} > gpurun_out/s8_step_variants.log 2>&1
cat gpurun_out/s8_step_variants.log
timeout -k 10 300 ncu --set full --import-source on --clock-control none -k regex:gemm2_kernel -c 1 -s 2 \
  -o gpurun_out/r02_gemm2_norm_rope python tools/gemm_one.py 8704 3072 3072 norm_rope > gpurun_out/s8_ncu_norm_rope.log 2>&1
tail -3 gpurun_out/s8_ncu_norm_rope.log
