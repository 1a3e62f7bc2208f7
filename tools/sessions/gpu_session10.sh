#!/usr/bin/env bash
# ncu --set full of the 1-CTA GEMM at REGION-step shapes (evidence for the "narrow tiles are L2-fill co-bound" note)
set -u
O=gpurun_out; T=${1:-r01s13}; mkdir -p $O
timeout -s INT 300 ncu --set full --clock-control none --import-source on -f -o /tmp/${T}_gemm1 -k regex:'gemm_kernel' -s 9 -c 8 \
    python tools/profile_step.py --blocks 1 1 --full 1 --region 1 --edited 1064 > $O/${T}_ncu_gemm1.log 2>&1
python tools/ncu_extract.py /tmp/${T}_gemm1.ncu-rep > $O/${T}_prof_gemm1_summary.csv 2>> $O/${T}_ncu_gemm1.log
tail -3 $O/${T}_ncu_gemm1.log; cut -c1-200 $O/${T}_prof_gemm1_summary.csv | head -12
