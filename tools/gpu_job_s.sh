#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py tests/test_gpu_sizes.py -x -q -s 2>&1 | grep -v "^$" | tail -6
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/s_bench.json 2> $O/s_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/s_bench.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['setup_ms'])
P
