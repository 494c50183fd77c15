#!/bin/bash
# 2-GPU pass: sharded-path parity test + scaling bench at N=1,2 for the reduce modes
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi_rank.py -m gpu -q -x -s > $O/h_pytest_mgpu.log 2>&1; tail -30 $O/h_pytest_mgpu.log
for M in 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --reduce-mode $M > $O/h_bench_n2_m$M.json 2> $O/h_bench_n2_m$M.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/h_bench_n2_m$M.json'))
    print('mode $M', {k:d[k] for k in ('value','ms_per_step','phases_ms','edges_linearized_per_s')}, d['e2e']['value'])
except Exception as e:
    print('mode $M failed', e); print(open('gpurun_out/h_bench_n2_m$M.err').read()[-1500:])
PY
done
