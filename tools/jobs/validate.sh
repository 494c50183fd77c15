#!/bin/bash
# GPU box: validation pass after a kernel change -- the whole GPU suite, smoke, the headline bench line
set -u
R=${1:-r02}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu_$R.log 2>&1; tail -9 $O/pytest_gpu_$R.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err; tail -2 $O/bench_$R.err
python - <<P
import json
d=json.loads(open('$O/bench_$R.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phases_ms'], d['pcg_iterations'], d['e2e']['value'], d['roofline']['frac'], d['roofline_linearize']['frac'], d['clocks'])
P
