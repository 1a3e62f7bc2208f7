#!/usr/bin/env bash
# validation of N-fast tile order, packed-rounding LayerNorm, 4-rows-per-warp GEMV: fast tests, isolated GEMMs, bench
set -u
O=gpurun_out; T=${1:-r01s8}; mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_flux_fullimage_gpu.py 2>&1 | tail -6) > $O/${T}_tests_fast.log
RGE_RASTER=m timeout 200 python tools/gemm_bench.py > $O/${T}_gemm_bench_mfast.log 2>&1
timeout 200 python tools/gemm_bench.py > $O/${T}_gemm_bench_auto.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
RGE_RASTER=m timeout 200 python tools/step_times.py > $O/${T}_step_times_mfast.log 2>&1
timeout 200 python tools/step_times.py > $O/${T}_step_times.log 2>&1
timeout 100 python tools/hbm_ops_bench.py > $O/${T}_hbm_ops.log 2>&1
tail -2 $O/${T}_tests_fast.log; cat $O/${T}_gemm_bench_mfast.log $O/${T}_gemm_bench_auto.log; cut -c1-250 $O/${T}_bench.json; tail -4 $O/${T}_step_times_mfast.log; tail -4 $O/${T}_step_times.log; tail -3 $O/${T}_hbm_ops.log
