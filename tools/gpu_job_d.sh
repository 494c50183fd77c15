#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pcg or precond" > $O/d_pytest.log 2>&1; tail -5 $O/d_pytest.log
for R in 1 8; do
timeout 300 python tools/prof_solve_steps.py $R > $O/d_steps_r$R.log 2>&1; tail -12 $O/d_steps_r$R.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/d_launches.csv python tools/prof_solve_steps.py 1 > $O/d_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/d_launches.csv',errors='ignore')) if len(r)>5]
h=rows[0]; kn,mv,mu=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
for r in rows[1:40]:
    if r[mv]: print(r[kn].split("(")[0][:60], float(r[mv].replace(",",""))*(1e-3 if r[mu] in("ns","nsecond") else 1.0))
PY
