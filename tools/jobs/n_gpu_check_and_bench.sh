#!/bin/bash
# N-GPU box: multi-rank parity check (all reduce modes) and the default bench line at N ranks
set -u
N=${1:-4}
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_check.py 20000 5000 200000 2>&1 | grep -v Warning | grep "rank 0 mode\|MGPU\|mismatch\|rror" | tail -10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_r02.json 2> $O/bench_n${N}_r02.err
python - $N <<'PY'
import json, sys
try:
    d=json.load(open('gpurun_out/bench_n%s_r02.json' % sys.argv[1]))
    print({k:d[k] for k in ('value','ms_per_step','phases_ms','edges_linearized_per_s')}, d['e2e']['value'], d['config']['parallelism'][:60], d['roofline_combine']['frac'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_n%s_r02.err' % sys.argv[1]).read()[-1500:])
PY
