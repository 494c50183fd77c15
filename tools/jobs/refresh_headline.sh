#!/bin/bash
# GPU box: refresh of the headline bench line and the launch list after late changes (the full pass is tools/collect_profiles.sh)
set -u
R=${1:-r02}
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err; tail -c 400 $O/bench_$R.json; tail -2 $O/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$R.log 2>&1
python tools/launch_list.py $O/launches_$R.csv > $O/launches_$R.txt 2>&1; tail -22 $O/launches_$R.txt
