#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > $O/f_pytest.log 2>&1; tail -12 $O/f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/f_bench.json 2> $O/f_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','pcg_iterations','edges_linearized_per_s')}); print(d['roofline_linearize']); print(d['e2e'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/f_ncu.log 2>&1
python tools/launch_list.py $O/f_launches.csv
