#!/bin/bash
# GPU box: last pass of a round -- validate.sh (GPU suite, smoke, headline bench line), then the other BASELINE workloads' bench lines
set -u
R=${1:-r02}
O=gpurun_out
bash tools/jobs/validate.sh $R
for W in full mini synth-100k batch-4096; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 5 > $O/bench_${W}_$R.json 2> $O/bench_${W}_$R.err
  python - <<P
import json
d=json.loads(open('$O/bench_${W}_$R.json').read().strip().splitlines()[-1])
print('$W', d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['clocks'].get('samples'), d.get('cpu_baseline',{}).get('value'))
P
done
