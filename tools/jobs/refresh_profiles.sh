#!/bin/bash
# GPU box: profile refresh after a kernel change -- launch list of the bench command and full captures of the kernels the bench reports a roofline for
set -u
R=${1:-r02}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$R.log 2>&1
python tools/launch_list.py $O/launches_$R.csv > $O/launches_$R.txt 2>&1; tail -16 $O/launches_$R.txt
ncu --set full --import-source on --clock-control none -k regex:"k_linearize_bearing_persistent|k_linearize_odometry" -s 6 -c 2 -f -o $O/prof_lin_$R python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_lin_$R.log 2>&1
python tools/ncu_summary.py $O/prof_lin_$R.ncu-rep 25 > $O/ncu_lin_$R.txt 2>&1; grep -n "Duration\|Executed Ipc Active\|dram__bytes\|stall mix" $O/ncu_lin_$R.txt
