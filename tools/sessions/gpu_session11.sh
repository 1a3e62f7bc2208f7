#!/usr/bin/env bash
# A/B of RGE_GROUP_QKV (q / k / v of one stream as one grouped launch inside the fan-out)
set -u
O=gpurun_out; T=${1:-r01s14}; mkdir -p $O
(timeout 200 python -m pytest tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -5) > $O/${T}_tests.log
RGE_GROUP_QKV=1 timeout 150 python tools/step_times.py > $O/${T}_step_times_group_qkv.log 2>&1
RGE_GROUP_QKV=0 timeout 150 python tools/step_times.py > $O/${T}_step_times_fanout.log 2>&1
tail -2 $O/${T}_tests.log; tail -4 $O/${T}_step_times_group_qkv.log; tail -4 $O/${T}_step_times_fanout.log
