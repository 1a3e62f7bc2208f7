#!/bin/bash
# Round 2, GPU session 15: attention tests after the merge-grid / default-offload change, the other families at full size
# through the CLI with the final code, step times.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_flux_fullsize_gpu.py \
  tests/test_qwen_fullsize_gpu.py > gpurun_out/s15_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s15_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s15_tests.log | tail -6
for fam in Step1X-Edit Step1X-Edit-v1p2 Qwen-Image; do
  timeout 300 python -m regione_b200.cli $fam --use_regione --erosion_dilation --model_path synthetic \
    --image_path assets/data.jsonl --output_dir /tmp/cli_$fam > gpurun_out/s15_cli_$fam.log 2>&1
  echo "== $fam"; grep -h "Time consuming" gpurun_out/s15_cli_$fam.log | tail -4
done
timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4 | grep -v SKIP
