#!/bin/bash
set -u
O=gpurun_out
for v in NO_OFFDIAG NO_CHAIN NO_LROWS; do
echo "== $v"
BOS_LIB_PATH=tools/_variants/libbos_b200_$v.so timeout 200 python tools/prof_solve.py 200000 50000 2000000 60 0 0 2>&1 | tail -2
done
echo "== baseline"
timeout 200 python tools/prof_solve.py 200000 50000 2000000 60 0 0 2>&1 | tail -2
