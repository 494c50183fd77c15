"""GPU box: dense Cholesky vs PCG solve time over problem sizes -> where BOS_SOLVER_AUTO should switch (dense_max_dim)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_problem, golden_problem, load_golden
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import xyt_to_xycs

rows = []
cases = [("mini", None), ("full", None)] + [("synth", n) for n in (100, 200, 400, 800, 1500, 3000, 6000, 10000)]
for name, n in cases:
    if n is None:
        g = load_golden(name); pr = golden_problem(g); P0 = g["poses_xycs"]
    else:
        w, pr = synth_problem(n, max(n // 5, 4), 10 * n, seed=0xB0500003); P0 = xyt_to_xycs(w["poses_init"])
    res = {"case": name if n is None else "synth-%d" % n, "n": 3 * pr.NP}
    for sname, solver in (("dense", capi.SOLVER_DENSE_CHOLESKY), ("sparse", capi.SOLVER_SPARSE_CHOLESKY), ("pcg", capi.SOLVER_PCG)):
        if sname == "dense" and 3 * pr.NP > 20000:
            continue
        ctx = capi.Context(solver=solver, pcg_rtol=1e-8, pcg_max_iters=20000)
        pr.upload(ctx); ctx.set_state(P0, None); ctx.triangulate()
        ms = []
        for it in range(6):
            s = ctx.step()
            if it >= 2:
                ms.append(s.ms_solve)
        res[sname + "_ms"] = float(np.mean(ms)); res[sname + "_status"] = int(s.solver_status)
        if sname == "pcg":
            res["pcg_iterations"] = int(s.pcg_iterations)
        ctx.close()
    rows.append(res)
    print(json.dumps(res), flush=True)
