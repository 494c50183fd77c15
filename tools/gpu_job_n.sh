#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_ref_build.py -q 2>&1 | tail -40 > $O/n_ref_tests.log
tail -5 $O/n_ref_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
