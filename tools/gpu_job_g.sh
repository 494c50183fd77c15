#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:"k_linearize_bearing_persistent|k_linearize_odometry|k_hb_init" -s 9 -c 3 -f -o $O/g_prof_lin python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/g_ncu.log 2>&1
python tools/ncu_summary.py $O/g_prof_lin.ncu-rep src > $O/g_ncu_lin.txt 2>&1; head -150 $O/g_ncu_lin.txt
