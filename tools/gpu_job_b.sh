#!/bin/bash
# round-2 GPU pass B: coarse space with several nodes per chunk (banded factorisation, lagged refresh)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -m gpu -q -x -k "pcg or precond or synth or large or converges or trajectory or smoke" --durations=8 > $O/b_pytest.log 2>&1; tail -25 $O/b_pytest.log
for R in 1 4; do
BOS_COARSE_REFRESH=$R timeout 300 python tools/prof_solve_steps.py $R > $O/b_steps_r$R.log 2>&1; tail -14 $O/b_steps_r$R.log
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/b_bench.json 2> $O/b_bench.err; tail -c 1800 $O/b_bench.json; tail -3 $O/b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/b_ncu.log 2>&1
python tools/launch_list.py $O/b_launches.csv > $O/b_launches.txt 2>&1; cat $O/b_launches.txt
