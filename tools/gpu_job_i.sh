#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sweep_solvers.py > $O/i_sweep.jsonl 2> $O/i_sweep.err; cat $O/i_sweep.jsonl; tail -3 $O/i_sweep.err
