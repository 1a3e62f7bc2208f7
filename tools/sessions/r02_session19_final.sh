#!/bin/bash
# Round 2, last GPU session: whole `pytest -m gpu` suite, smoke(), default bench line and step times on the final code.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/s19_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s19_tests.log
grep -E "passed|failed|FAILED|rc=|noise floor|true-CFG" gpurun_out/s19_tests.log | cut -c1-260 | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s19_smoke.log 2>&1; tail -2 gpurun_out/s19_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; tail -c 900 gpurun_out/s19_bench.json
timeout -k 10 200 python tools/step_times.py > gpurun_out/s19_step_times.log 2>&1; tail -4 gpurun_out/s19_step_times.log | grep -v SKIP
