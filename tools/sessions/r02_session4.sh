#!/bin/bash
# Round 2, GPU session 4: whole GPU suite after the refactors, compute-sanitizer runs, ncu traffic capture of the
# CTA-pair GEMMs of a FULL step, ncu launch list of one image, bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/s4_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_tests.log
grep -E "passed|failed|FAILED|rc=|noise floor|true-CFG|bit-identical" gpurun_out/s4_tests.log | cut -c1-330 | tail -20
for knobs in "RGE_NVTX=0" "RGE_NO_FANOUT=1" "RGE_GROUPED=1"; do
  for tool in memcheck racecheck; do
    env $knobs timeout -k 10 420 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 \
      python -m pytest -q -p no:cacheprovider -m gpu tests/test_flux_parity_gpu.py::test_tiny_flux_default_schedule \
      > "gpurun_out/s4_sanitizer_${tool}_${knobs}.log" 2>&1
    echo "$tool $knobs rc=$?" | tee -a gpurun_out/s4_sanitizer_summary.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "gpurun_out/s4_sanitizer_${tool}_${knobs}.log" | tail -3 | tee -a gpurun_out/s4_sanitizer_summary.log
  done
done
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_kernel \
  -o gpurun_out/r02_gemm_full_step python tools/profile_step.py --blocks 1 1 --full 3 --region 0 --profiler-range \
  > gpurun_out/s4_ncu_gemm.log 2>&1; tail -3 gpurun_out/s4_ncu_gemm.log
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_image.csv python bench.py --steps 1 --warmup 3 --profiler-range --no-cpu-baseline \
  --no-reference-gpu > gpurun_out/s4_ncu_launches.log 2>&1; wc -l gpurun_out/r02_launches_image.csv
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; tail -c 1500 gpurun_out/s4_bench.json
