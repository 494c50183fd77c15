#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py tests/test_ref_build.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/t_bench.json 2> $O/t_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/t_bench.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['roofline_linearize']['frac'], d['roofline']['frac'])
P
