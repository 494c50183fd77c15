#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/c_launches.csv python bench.py --steps 1 --warmup 4 --no-cpu-baseline > $O/c_ncu.log 2>&1
python tools/launch_list.py $O/c_launches.csv > $O/c_launches.txt 2>&1; cat $O/c_launches.txt
