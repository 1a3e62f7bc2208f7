#!/usr/bin/env bash
# One GPU-box session: tests, reference-equivalent GPU arm, rho sweep, HBM-op microbench, ncu launch list + full sets.
set -u
O=gpurun_out; T=${1:-r01s3}; mkdir -p $O
(python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $O/${T}_tests.log
timeout 900 python bench.py --impl reference_gpu --warmup 2 --steps 2 > $O/${T}_reference_gpu.json 2> $O/${T}_reference_gpu.err
timeout 600 python tools/rho_sweep.py > $O/${T}_rho_sweep.log 2>&1
timeout 300 python tools/hbm_ops_bench.py > $O/${T}_hbm_ops.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profiler-range > $O/${T}_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -f -o /tmp/${T}_full \
    python tools/profile_step.py --blocks 1 1 --full 1 --region 1 > $O/${T}_full_run.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > $O/${T}_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -f -o /tmp/${T}_hbm \
    -k regex:'arp_|euler|move_rows|cfg_|ln_modulate|morph' python tools/hbm_ops_bench.py --once --images 64 > $O/${T}_hbm_run.log 2>&1
ncu -i /tmp/${T}_hbm.ncu-rep --page raw --csv > $O/${T}_hbm_raw.csv 2>/dev/null
ls -la /tmp/${T}_*.ncu-rep
for f in /tmp/${T}_full.ncu-rep /tmp/${T}_hbm.ncu-rep; do [ $(stat -c %s $f) -lt 25000000 ] && cp $f $O/; done
tail -3 $O/${T}_tests.log; cat $O/${T}_reference_gpu.json; cat $O/${T}_rho_sweep.log | tail -12; cat $O/${T}_hbm_ops.log | tail -20; wc -l $O/${T}_launches.csv $O/${T}_full_raw.csv $O/${T}_hbm_raw.csv
