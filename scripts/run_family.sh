#!/usr/bin/env bash
# Counterpart of the reference's script/<Family>.sh (demo + evaluation invocations) over regione_b200.cli.
#   scripts/run_family.sh <FluxKontext|Step1X-Edit|Step1X-Edit-v1p2|Qwen-Image|Qwen-Image-Edit-2509> [demo|eval] [model_path]
# model_path defaults to `synthetic` (no weights exist offline); pass a diffusers checkpoint directory to use a real model.
set -euo pipefail
family=${1:?family}; mode=${2:-demo}; model=${3:-synthetic}
case "$family" in
  FluxKontext)            thr=0.93; demo_cache=0.01; eval_cache=0.04; g=2.5; bench=Kontext-Bench ;;
  Step1X-Edit|Step1X-Edit-v1p2) thr=0.88; demo_cache=0.02; eval_cache=0.02; g=6.0; bench=GEdit-Bench ;;
  Qwen-Image|Qwen-Image-Edit-2509) thr=0.80; demo_cache=0.03; eval_cache=0.03; g=4.0; bench=GEdit-Bench ;;
  *) echo "unknown family $family" >&2; exit 2 ;;
esac
common=(--model_path "$model" --num_inference_steps 28 --use_regione --warmup_step 6 --post_step 2 --refresh_step "16"
        --threshold "$thr" --erosion_dilation --guidance_scale "$g" --seed 110 --device cuda)
if [ "$mode" = demo ]; then
  python -m regione_b200.cli "$family" "${common[@]}" --cache_threshold "$demo_cache" \
      --image_path "${IMAGE_PATH:-assets/data.jsonl}" --output_dir "result/$family/Demo/RegionE"
else
  python -m regione_b200.cli "$family" "${common[@]}" --cache_threshold "$eval_cache" --evaluation \
      --image_path "${IMAGE_PATH:-data/Processed/$bench}" --output_dir "result/$family/RegionE"
fi
