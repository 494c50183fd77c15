#!/bin/bash
set -u
O=gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > $O/k_bench.json 2> $O/k_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/k_bench.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['pcg_iterations'], d['e2e']['value'], d['roofline']['frac'], d['roofline_linearize']['frac'], d['pcg_iterations_per_step'])
P
timeout 600 python -m pytest tests -m gpu -x -q -k "pcg or lagged or refresh or coarse or run_to_run or precond" 2>&1 | tail -3
