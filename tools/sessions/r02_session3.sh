#!/bin/bash
# Round 2, GPU session 3: fixed pipe micro-benchmarks, ncu source-level profiles of the attention kernel (whole-row and
# pipelined-halves variants), attention / GEMM variant benchmarks, launch-schedule variants of the REGION / FULL steps.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I regione_b200/csrc -o /tmp/pipes tools/microbench/pipes.cu \
  && timeout 60 /tmp/pipes > gpurun_out/s3_pipes.log 2>&1; cat gpurun_out/s3_pipes.log
timeout 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py \
  tests/test_partially_linear_vs_triton_gpu.py > gpurun_out/s3_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s3_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s3_tests.log | tail -12
timeout 300 python tools/attn_bench.py > gpurun_out/s3_attn_bench.log 2>&1; cat gpurun_out/s3_attn_bench.log
for pipe in 0 1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 1 -c 1 \
    -o gpurun_out/r02_attn_pipe$pipe python tools/attn_one.py 8704 8704 0 $pipe > gpurun_out/s3_ncu_attn$pipe.log 2>&1
  tail -2 gpurun_out/s3_ncu_attn$pipe.log
done
timeout 500 python tools/gemm_bench.py --quick > gpurun_out/s3_gemm_bench.log 2>&1; tail -9 gpurun_out/s3_gemm_bench.log
run_steps() { echo "== $1"; env $1 timeout 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_NVTX=0"
  run_steps "RGE_GROUP_QKV=1"
  run_steps "RGE_GROUP_QKV=2"
  run_steps "RGE_GROUP_QKV=2 RGE_GEMM3=1"
  run_steps "RGE_GROUPED=1 RGE_GEMM3=1"
  run_steps "RGE_ATTN_PIPE=1"
} > gpurun_out/s3_step_variants.log 2>&1
cat gpurun_out/s3_step_variants.log
