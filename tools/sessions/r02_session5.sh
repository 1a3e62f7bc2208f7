#!/bin/bash
# Round 2, GPU session 5: the decoupled-pipeline attention kernel (attention64.cu) against attention.cu: parity tests of
# both, isolated throughput, step times, ncu source-level profile.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py -k "attention or pack_unpack" \
  > gpurun_out/s5_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s5_tests.log | tail -12
timeout -k 10 300 python tools/attn_bench.py > gpurun_out/s5_attn_bench.log 2>&1; cat gpurun_out/s5_attn_bench.log | cut -c1-900
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{
  run_steps "RGE_ATTN_KERNEL=0"
  run_steps "RGE_ATTN_KERNEL=1"
} > gpurun_out/s5_step_variants.log 2>&1
cat gpurun_out/s5_step_variants.log
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attention64 -s 1 -c 1 \
  -o gpurun_out/r02_attn64 python tools/attn_one.py 8704 8704 1 > gpurun_out/s5_ncu_attn64.log 2>&1
tail -2 gpurun_out/s5_ncu_attn64.log
