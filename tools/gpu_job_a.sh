#!/bin/bash
# round-2 GPU pass A: parity tests (incl. the BASELINE sizes), PCG phase timing (diagnostic build), FP64 GEMM peak, bench
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=20 > $O/a_pytest_gpu.log 2>&1; tail -40 $O/a_pytest_gpu.log
BOS_LIB_PATH=tools/_variants/libbos_b200_timing.so timeout 300 python tools/prof_solve.py 200000 50000 2000000 20000 0 0 > $O/a_pcg_timing_p0.log 2>&1; tail -12 $O/a_pcg_timing_p0.log
BOS_LIB_PATH=tools/_variants/libbos_b200_timing.so timeout 300 python tools/prof_solve.py 200000 50000 2000000 300 0 1 > $O/a_pcg_timing_p1.log 2>&1; tail -8 $O/a_pcg_timing_p1.log
timeout 300 python tools/peaks_fp64.py > $O/a_fp64_peak.json 2>&1; cat $O/a_fp64_peak.json
timeout 600 python bench.py --steps 10 --warmup 3 > $O/a_bench.json 2> $O/a_bench.err; tail -c 1500 $O/a_bench.json; tail -3 $O/a_bench.err
