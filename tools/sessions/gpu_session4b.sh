#!/usr/bin/env bash
# final round-1 numbers: all GPU tests, bench (both arms), per-step times, rho sweep, full-size runs of the other families through the CLI
set -u
O=gpurun_out; T=${1:-r01s7}; mkdir -p $O
(timeout 400 python -m pytest tests -m gpu -x -q --durations=4 2>&1 | tail -12) > $O/${T}_tests.log
timeout 300 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
timeout 200 python tools/step_times.py > $O/${T}_step_times.log 2>&1
timeout 400 python tools/rho_sweep.py > $O/${T}_rho_sweep.log 2>&1
for fam in Step1X-Edit Step1X-Edit-v1p2 Qwen-Image; do
  timeout 300 python -m regione_b200.cli $fam --use_regione --erosion_dilation --model_path synthetic --image_path assets/data.jsonl --output_dir /tmp/cli_$fam > $O/${T}_cli_$fam.log 2>&1
done
tail -3 $O/${T}_tests.log; cat $O/${T}_bench.json | cut -c1-300; cat $O/${T}_bench_reference.json | cut -c1-300; tail -4 $O/${T}_step_times.log; tail -12 $O/${T}_rho_sweep.log; grep -h "Time consuming" $O/${T}_cli_*.log | head -20
