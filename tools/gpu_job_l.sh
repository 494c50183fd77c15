#!/bin/bash
set -u
O=gpurun_out
ncu --set full --import-source on --clock-control none -k k_pcg_fused -s 3 -c 1 -f -o $O/prof_pcg_r02b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_pcg_r02b.log 2>&1
python tools/ncu_summary.py $O/prof_pcg_r02b.ncu-rep 25 > $O/ncu_pcg_r02b.txt 2>&1
tail -5 $O/ncu_pcg_r02b.log
