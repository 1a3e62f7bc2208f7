#!/usr/bin/env bash
# grouped-GEMM validation: tests (fast ones first), bench, step times grouped / ungrouped, region timeline
set -u
O=gpurun_out; T=${1:-r01s5}; mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_flux_fullimage_gpu.py 2>&1 | tail -15) > $O/${T}_tests_fast.log
timeout 300 python bench.py --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
RGE_GROUPED=0 timeout 200 python tools/step_times.py > $O/${T}_step_times_ungrouped.log 2>&1
RGE_GROUPED=1 timeout 200 python tools/step_times.py > $O/${T}_step_times_grouped.log 2>&1
RGE_GROUPED=1 RGE_FILL_ATTN_TAIL=0 timeout 200 python tools/step_times.py > $O/${T}_step_times_grouped_nofill.log 2>&1
timeout 120 python tools/timeline.py > $O/${T}_timeline_grouped.log 2>&1
(timeout 300 python -m pytest tests/test_flux_fullimage_gpu.py -m gpu -x -q -s 2>&1 | tail -6) > $O/${T}_tests_fullimage.log
tail -3 $O/${T}_tests_fast.log; cat $O/${T}_bench.json | cut -c1-400; tail -4 $O/${T}_step_times_ungrouped.log; tail -4 $O/${T}_step_times_grouped.log; tail -4 $O/${T}_step_times_grouped_nofill.log; tail -4 $O/${T}_tests_fullimage.log
