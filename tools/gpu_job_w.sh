#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py tests/test_ref_build.py -x -q -k "dense or sparse or skyline or solve or trajectory or cholesky or reference" 2>&1 | tail -3
for sv in dense sparse; do
for ng in 0 1; do
if [ $ng = 1 ]; then export BOS_NO_GRAPH=1; else unset BOS_NO_GRAPH; fi
timeout 300 python bench.py --workload synth-100k --solver $sv --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$sv no_graph=$ng', d['value'], d['phases_ms'])"
done; done
unset BOS_NO_GRAPH
timeout 300 python bench.py --workload full --solver dense --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('full dense graph', d['value'], d['phases_ms'])"
BOS_NO_GRAPH=1 timeout 300 python bench.py --workload full --solver dense --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('full dense eager', d['value'], d['phases_ms'])"
