#!/bin/bash
# GPU box: H, b build variants (tools/_variants/libbos_b200_<V>.so, built with -D... by build.build(extra_flags=, out=, objdir=)) timed beside the default library,
# then the H, b parity tests on every variant
set -u
O=gpurun_out
echo "== default"; timeout 200 python tools/prof_lin.py 200000 50000 2000000 40 2>&1 | tail -1
for v in "$@"; do
  echo "== $v"; BOS_LIB_PATH=$PWD/tools/_variants/libbos_b200_$v.so timeout 200 python tools/prof_lin.py 200000 50000 2000000 40 2>&1 | tail -1
done
echo "== default again"; timeout 200 python tools/prof_lin.py 200000 50000 2000000 40 2>&1 | tail -1
for v in "$@"; do
  echo "== parity $v"; BOS_LIB_PATH=$PWD/tools/_variants/libbos_b200_$v.so timeout 400 python -m pytest tests -m gpu -x -q -k "H_b or edge_free or duplicate_blocks or odometry_only or robust" 2>&1 | tail -3
done
