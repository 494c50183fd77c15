#!/bin/bash
set -u
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_check.py 3000 700 30000 2>&1 | grep -v Warning | grep "mode 5\|mode 3\|MGPU\|mismatch\|Error\|error" | tail -12
bash tools/gpu_job_p.sh 2>&1 | tail -4
