#!/bin/bash
# 2-GPU box: the multi-rank parity tests and the N = 2 bench line of the round
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi_rank.py -m gpu -x -q > $O/pytest_gpu_2ranks_r02.log 2>&1; tail -3 $O/pytest_gpu_2ranks_r02.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_r02.json 2> $O/bench_n2_r02.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_r02.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','edges_linearized_per_s')}, d['e2e']['value'], d['config']['parallelism'])
PY
