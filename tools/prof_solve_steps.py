"""GPU box: 12 GN steps of synth-2M with a given coarse refresh period: per-step CG iterations and phase times."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import xyt_to_xycs
refresh = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w, pr, _ = bench.make_world("synth-2M")
ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=1e-8, pcg_max_iters=20000, pcg_coarse_refresh=refresh, pcg_coarse_nodes=nodes)
pr.upload(ctx)
ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
ctx.triangulate()
for it in range(12):
    s = ctx.step()
    print("step %2d chi2 %.6e pcg %3d status %d precond %d  ms lin %.3f solve %.3f (%.1f us/iter) upd %.3f launches %d" % (
        it, s.chi2_bearing + s.chi2_odometry, s.pcg_iterations, s.solver_status, s.precond_used, s.ms_linearize, s.ms_solve,
        1e3 * s.ms_solve / max(s.pcg_iterations, 1), s.ms_update, s.gpu_launches), flush=True)
