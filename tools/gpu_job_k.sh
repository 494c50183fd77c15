#!/bin/bash
set -u
BOS_LIB_PATH=tools/_variants/libbos_b200_timing.so timeout 200 python tools/prof_solve.py 200000 50000 2000000 60 0 0 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/k_launches.csv python tools/prof_solve.py 200000 50000 2000000 60 0 0 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/k_launches.csv 2>&1 | tail -14
