#!/bin/bash
# Round 2, GPU session 6: warp-uniform issue (elect_one) in every TMA producer / MMA issuer loop - whole suite, isolated
# attention and GEMM throughput, step times with both attention kernels, bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 500 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s6_tests_fast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s6_tests_fast.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s6_tests_fast.log | tail -12
timeout -k 10 300 python tools/attn_bench.py > gpurun_out/s6_attn_bench.log 2>&1; cat gpurun_out/s6_attn_bench.log | cut -c1-900
timeout -k 10 500 python tools/gemm_bench.py --quick > gpurun_out/s6_gemm_bench.log 2>&1; cat gpurun_out/s6_gemm_bench.log | cut -c1-700
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_ATTN_KERNEL=0"
  run_steps "RGE_ATTN_KERNEL=1"
  run_steps "RGE_ATTN_KERNEL=0 RGE_GROUP_QKV=0"
  run_steps "RGE_ATTN_KERNEL=0 RGE_GROUP_QKV=2"
} > gpurun_out/s6_step_variants.log 2>&1
cat gpurun_out/s6_step_variants.log
timeout -k 10 1200 python -m pytest -q -rA -p no:cacheprovider -m gpu tests --deselect tests/test_kernels_gpu.py \
  --deselect tests/test_trim_last_gpu.py --deselect tests/test_flux_parity_gpu.py --deselect tests/test_flux_fullsize_gpu.py \
  > gpurun_out/s6_tests_rest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s6_tests_rest.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s6_tests_rest.log | tail -12
