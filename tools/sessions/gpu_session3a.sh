#!/usr/bin/env bash
# tests + bench + timelines with / without the attention-tail fill + HBM-op microbench (no ncu): ~6 min
set -u
O=gpurun_out; T=${1:-r01s4}; mkdir -p $O
(timeout 500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25) > $O/${T}_tests.log
timeout 300 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
RGE_FILL_ATTN_TAIL=0 timeout 200 python tools/step_times.py > $O/${T}_step_times_nofill.log 2>&1
RGE_FILL_ATTN_TAIL=1 timeout 200 python tools/step_times.py > $O/${T}_step_times_fill.log 2>&1
RGE_FILL_ATTN_TAIL=0 timeout 120 python tools/timeline.py > $O/${T}_timeline_nofill.log 2>&1
RGE_FILL_ATTN_TAIL=1 timeout 120 python tools/timeline.py > $O/${T}_timeline_fill.log 2>&1
timeout 120 python tools/hbm_ops_bench.py > $O/${T}_hbm_ops.log 2>&1
tail -4 $O/${T}_tests.log; cat $O/${T}_bench.json; tail -5 $O/${T}_step_times_nofill.log; tail -5 $O/${T}_step_times_fill.log; cat $O/${T}_hbm_ops.log
