#!/usr/bin/env bash
# ncu captures of the dominant kernels (3 launches each, --set full) + a metrics-only pass over the HBM-bound kernels
set -u
O=gpurun_out; T=${1:-r01s6}; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s INT 330 $NCU -o /tmp/${T}_gemm2 -k regex:gemm2_kernel -s 4 -c 3 python tools/profile_step.py --blocks 1 1 --full 1 --region 0 > $O/${T}_ncu_gemm2.log 2>&1
python tools/ncu_extract.py /tmp/${T}_gemm2.ncu-rep > $O/${T}_prof_gemm2_summary.csv 2>> $O/${T}_ncu_gemm2.log
timeout -s INT 330 $NCU -o /tmp/${T}_attn -k regex:attention_kernel -c 3 python tools/profile_step.py --blocks 1 1 --full 1 --region 1 > $O/${T}_ncu_attn.log 2>&1
python tools/ncu_extract.py /tmp/${T}_attn.ncu-rep > $O/${T}_prof_attn_summary.csv 2>> $O/${T}_ncu_attn.log
timeout -s INT 300 $NCU -o /tmp/${T}_small -k regex:'ln_modulate|gemv_batch|gemm_kernel' -s 4 -c 4 python tools/profile_step.py --blocks 1 1 --full 1 --region 0 > $O/${T}_ncu_small.log 2>&1
python tools/ncu_extract.py /tmp/${T}_small.ncu-rep > $O/${T}_prof_small_summary.csv 2>> $O/${T}_ncu_small.log
timeout -s INT 200 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    -k regex:'arp_|euler|move_rows|cfg_|morph|pack_lat' --csv --log-file $O/${T}_hbm_ncu.csv python tools/hbm_ops_bench.py --once --images 64 > $O/${T}_ncu_hbm.log 2>&1
ls -la /tmp/${T}_*.ncu-rep
for f in /tmp/${T}_gemm2.ncu-rep /tmp/${T}_attn.ncu-rep /tmp/${T}_small.ncu-rep; do [ -f $f ] && [ $(stat -c %s $f) -lt 15000000 ] && cp $f $O/; done
cat $O/${T}_prof_gemm2_summary.csv | cut -c1-600; tail -3 $O/${T}_ncu_attn.log
