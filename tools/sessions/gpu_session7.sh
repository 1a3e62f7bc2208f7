#!/usr/bin/env bash
# ncu launch list of one FULL + one REGION step at full depth (the last step of each kind is inside the profiler range)
set -u
O=gpurun_out; T=${1:-r01s10}; mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_flux_fullimage_gpu.py 2>&1 | tail -6) > $O/${T}_tests_fast.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches_steps.csv python tools/profile_step.py --full 2 --region 2 --edited 1064 --profiler-range > $O/${T}_launches_steps.log 2>&1
tail -2 $O/${T}_tests_fast.log; wc -l $O/${T}_launches_steps.csv; tail -5 $O/${T}_launches_steps.log
