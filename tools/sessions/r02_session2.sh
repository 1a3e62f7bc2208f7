#!/bin/bash
# Round 2, GPU session 2: pipe micro-benchmarks, ncu source-level profile of the attention kernel, the tests that changed
# since session 1 (gemm3, trimming, true-CFG, Triton thresholds, configs[2] floor), gemm3 benchmarks and step times.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I regione_b200/csrc -o /tmp/pipes tools/microbench/pipes.cu \
  && timeout 60 /tmp/pipes > gpurun_out/s2_pipes.log 2>&1; cat gpurun_out/s2_pipes.log
timeout 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_partially_linear_vs_triton_gpu.py tests/test_flux_parity_gpu.py > gpurun_out/s2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s2_tests.log | tail -12
timeout 500 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_qwen_config2_gpu.py > gpurun_out/s2_tests_config2.log 2>&1
grep -E "configs\[2\]|passed|failed" gpurun_out/s2_tests_config2.log | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 \
  -o gpurun_out/r02_attn_poly0 python tools/attn_one.py 8704 8704 0 > gpurun_out/s2_ncu_attn.log 2>&1; tail -2 gpurun_out/s2_ncu_attn.log
timeout 500 python tools/gemm_bench.py --quick > gpurun_out/s2_gemm_bench.log 2>&1; tail -30 gpurun_out/s2_gemm_bench.log
for mode in 0 1 2; do
  RGE_GEMM3=$mode timeout 200 python tools/step_times.py > gpurun_out/s2_step_times_gemm3_$mode.log 2>&1
  echo "RGE_GEMM3=$mode"; tail -4 gpurun_out/s2_step_times_gemm3_$mode.log
done
