#!/bin/bash
# 2-GPU box: N = 2 bench lines with the fused peer build (reduce_mode 4) and the NCCL ownership combine (3) beside it
set -u
O=gpurun_out
mkdir -p $O
for m in 5 3; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$m bench.py --gpus 2 --steps 20 --warmup 5 --reduce-mode $m > $O/bench_n2_mode${m}_r02.json 2> $O/bench_n2_mode${m}_r02.err
python - $m <<'PY'
import json, sys
try:
    d=json.load(open('gpurun_out/bench_n2_mode%s_r02.json' % sys.argv[1]))
    print({k:d[k] for k in ('value','ms_per_step','phases_ms','edges_linearized_per_s')}, d['e2e']['value'], d['roofline_combine'], d['chi2_last'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_n2_mode%s_r02.err' % sys.argv[1]).read()[-2000:])
PY
done
