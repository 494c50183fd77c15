#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for M in 3 2; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$M bench.py --gpus 2 --steps 10 --warmup 3 --reduce-mode $M > $O/h_bench_n2_m$M.json 2> $O/h_bench_n2_m$M.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/h_bench_n2_m$M.json'))
    print('mode $M', {k:d[k] for k in ('value','ms_per_step','phases_ms','edges_linearized_per_s')}, d['e2e']['value'])
except Exception as e:
    print('mode $M failed', e); print(open('gpurun_out/h_bench_n2_m$M.err').read()[-1500:])
PY
done
