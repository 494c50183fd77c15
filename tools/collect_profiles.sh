#!/bin/bash
# GPU box: the measurement pass of a round -- tests, bench (both arms), ncu launch list and full captures of the top kernels.
# Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into the committed files under profiles/.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$R.log 2>&1; tail -3 $O/pytest_gpu_$R.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2> $O/bench_ref_$R.err; tail -c 600 $O/bench_ref_$R.json
python bench.py --steps 5 --warmup 3 > $O/bench_$R.json 2> $O/bench_$R.err; tail -c 400 $O/bench_$R.json; tail -3 $O/bench_$R.err
# per-launch durations of the same command (cold-cache, serialised: only the kernels' SHARE of a step is comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$R.log 2>&1
# full captures of the two kernels the bench reports a roofline for, and of the pose kernel of the H, b build
ncu --set full --import-source on --clock-control none -k k_pcg_fused -s 3 -c 1 -f -o $O/prof_pcg_$R python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_pcg_$R.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_linearize_bearing_persistent|k_pose_finish" -s 6 -c 2 -f -o $O/prof_lin_$R python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_lin_$R.log 2>&1
# dense Cholesky (config 3): the DMMA trailing-update kernel of the first outer panels, and the solve time
python tools/prof_dense.py > $O/dense_$R.log 2>&1; tail -1 $O/dense_$R.log
ncu --set full --clock-control none -k k_syrk_big -s 2 -c 1 -f -o $O/prof_syrk_$R python tools/prof_dense.py > $O/ncu_syrk_$R.log 2>&1
# secondary lines: FP32 flavour of the same bench, dense workload, batched config 5
python bench.py --precision f32 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_f32_$R.json 2> $O/bench_f32_$R.err; tail -c 300 $O/bench_f32_$R.json
python bench.py --pcg-precond 1 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_bj_$R.json 2> $O/bench_bj_$R.err; tail -c 200 $O/bench_bj_$R.json
python bench.py --workload synth-100k --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_100k_$R.json 2> $O/bench_100k_$R.err
python tools/prof_batch.py 4096 > $O/batch_$R.log 2>&1; cat $O/batch_$R.log
ls -la $O | tail -12
