#!/usr/bin/env bash
# final check of the round's last commit: smoke(), every GPU test, the default bench line
set -u
O=gpurun_out; T=${1:-r01s11}; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
(timeout 500 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -8) > $O/${T}_tests.log
timeout 300 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
tail -2 $O/${T}_smoke.log; tail -3 $O/${T}_tests.log; cat $O/${T}_bench.json
