#!/bin/bash
# GPU box: the measurement pass of a round -- tests, bench (all workloads), ncu launch list and full captures of the top kernels.
# Everything lands in gpurun_out/; the files worth keeping are copied to profiles/ by hand (see profiles/README.md).
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu_$R.log 2>&1; tail -14 $O/pytest_gpu_$R.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err; tail -c 600 $O/bench_$R.json; tail -3 $O/bench_$R.err
for W in full mini synth-100k batch-4096; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 5 > $O/bench_${W}_$R.json 2> $O/bench_${W}_$R.err; tail -c 300 $O/bench_${W}_$R.json; echo
  timeout 300 python bench.py --impl reference --workload $W --steps 5 --warmup 2 > $O/bench_ref_${W}_$R.json 2> $O/bench_ref_${W}_$R.err; tail -c 200 $O/bench_ref_${W}_$R.json; echo
done
timeout 300 python bench.py --workload synth-100k --solver dense --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_synth-100k_dense_$R.json 2> $O/bench_synth-100k_dense_$R.err
timeout 300 python bench.py --pcg-precond 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_bj_$R.json 2> $O/bench_bj_$R.err; tail -c 200 $O/bench_bj_$R.json; echo
timeout 300 python bench.py --precision f32 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_f32_$R.json 2> $O/bench_f32_$R.err; tail -c 200 $O/bench_f32_$R.json; echo
# per-launch durations of the same command (cold-cache, serialised: only the kernels' SHARE of a step is comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$R.log 2>&1
python tools/launch_list.py $O/launches_$R.csv > $O/launches_$R.txt 2>&1; cat $O/launches_$R.txt
# full captures of the kernels the bench reports a roofline for
ncu --set full --import-source on --clock-control none -k k_pcg_fused -s 3 -c 1 -f -o $O/prof_pcg_$R python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_pcg_$R.log 2>&1
python tools/ncu_summary.py $O/prof_pcg_$R.ncu-rep 25 > $O/ncu_pcg_$R.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_linearize_bearing_persistent|k_linearize_odometry" -s 6 -c 2 -f -o $O/prof_lin_$R python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_lin_$R.log 2>&1
python tools/ncu_summary.py $O/prof_lin_$R.ncu-rep 25 > $O/ncu_lin_$R.txt 2>&1
python tools/peaks_fp64.py > $O/fp64_peak_$R.json 2>&1
ls -la $O | tail -30
