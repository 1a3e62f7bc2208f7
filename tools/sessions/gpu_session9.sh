#!/usr/bin/env bash
# check of the StepRun refactor: fast GPU tests (all launch-schedule variants), per-step times
set -u
O=gpurun_out; T=${1:-r01s12}; mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_flux_fullimage_gpu.py 2>&1 | tail -8) > $O/${T}_tests_fast.log
timeout 200 python tools/step_times.py > $O/${T}_step_times.log 2>&1
tail -3 $O/${T}_tests_fast.log; tail -4 $O/${T}_step_times.log
