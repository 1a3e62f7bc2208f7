#!/bin/bash
# Round 2, final GPU session: whole `pytest -m gpu` suite, smoke(), the default bench line (with the reference-equivalent
# GPU arm), `--impl reference`, isolated attention / GEMM throughput, rho sweep, ncu launch list of one FULL + one REGION
# step at full depth with the final code.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/s12_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s12_tests.log
grep -E "passed|failed|FAILED|rc=|noise floor|true-CFG|bit-identical" gpurun_out/s12_tests.log | cut -c1-300 | tail -16
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s12_smoke.log 2>&1; tail -2 gpurun_out/s12_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err; tail -c 1200 gpurun_out/s12_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s12_bench_reference.json 2> gpurun_out/s12_bench_reference.err; tail -c 600 gpurun_out/s12_bench_reference.json
timeout -k 10 300 python tools/attn_bench.py > gpurun_out/s12_attn_bench.log 2>&1; cut -c1-900 gpurun_out/s12_attn_bench.log
timeout -k 10 600 python tools/gemm_bench.py --quick > gpurun_out/s12_gemm_bench.log 2>&1; head -9 gpurun_out/s12_gemm_bench.log | cut -c1-500
timeout -k 10 400 python tools/rho_sweep.py > gpurun_out/s12_rho_sweep.log 2>&1; grep -v RegionEHelper gpurun_out/s12_rho_sweep.log
timeout -k 10 200 python tools/step_times.py > gpurun_out/s12_step_times.log 2>&1; tail -4 gpurun_out/s12_step_times.log | grep -v SKIP
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_steps_final.csv python tools/profile_step.py --full 2 --region 2 --edited 1064 \
  --profiler-range > gpurun_out/s12_ncu_launches.log 2>&1; wc -l gpurun_out/r02_launches_steps_final.csv
