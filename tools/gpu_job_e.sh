#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -m gpu -q -x -k "pcg or precond or odometry_only or large or duplicate" > $O/e_pytest.log 2>&1; tail -5 $O/e_pytest.log
timeout 300 python tools/prof_solve_steps.py 8 > $O/e_steps_r8.log 2>&1; tail -12 $O/e_steps_r8.log
BOS_LIB_PATH=tools/_variants/libbos_b200_timing.so timeout 300 python tools/prof_solve.py 200000 50000 2000000 20000 0 0 > $O/e_pcg_timing_p0.log 2>&1; tail -5 $O/e_pcg_timing_p0.log
