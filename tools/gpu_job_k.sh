#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -m gpu -x -q -k "sparse or solve_and_update or config3 or dense" -s 2>&1 | grep -E "passed|failed|^E  |dense .* ms|synth-100k" | head -20
timeout 600 python tools/sweep_solvers.py > $O/k_sweep.jsonl 2> $O/k_sweep.err; cat $O/k_sweep.jsonl; tail -3 $O/k_sweep.err
