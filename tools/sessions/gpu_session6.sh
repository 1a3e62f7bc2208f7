#!/usr/bin/env bash
# 2-GPU box: config 4 (one image stream per rank + NCCL mask all-gather) and config 5 (Step1X-Edit v1p2 data-parallel
# sweep through the CLI)
set -u
O=gpurun_out; T=${1:-r01s9}; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_bench_2gpu.json 2> $O/${T}_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    -m regione_b200.cli Step1X-Edit-v1p2 --use_regione --erosion_dilation --model_path synthetic --rho sweep \
    --image_path assets/sweep16.jsonl --output_dir /tmp/sweep_v1p2 > $O/${T}_dp_sweep_v1p2.log 2>&1
cut -c1-300 $O/${T}_bench_2gpu.json; grep -E "rank [0-9]+:|Time consuming" $O/${T}_dp_sweep_v1p2.log | tail -20
