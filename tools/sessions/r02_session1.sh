#!/bin/bash
# Round 2, GPU session 1: new parity tests + whole suite, attention / GEMM micro-benchmarks (variant sweeps), bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/s1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_tests.log
tail -5 gpurun_out/s1_tests.log
timeout 300 python tools/attn_bench.py > gpurun_out/s1_attn_bench.log 2>&1; tail -6 gpurun_out/s1_attn_bench.log
timeout 400 python tools/gemm_bench.py > gpurun_out/s1_gemm_bench.log 2>&1; tail -20 gpurun_out/s1_gemm_bench.log
timeout 200 python tools/step_times.py > gpurun_out/s1_step_times.log 2>&1; tail -6 gpurun_out/s1_step_times.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 3000 gpurun_out/s1_bench.json
