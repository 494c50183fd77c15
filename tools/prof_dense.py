"""GPU box: one dense-Cholesky solve of the synth-100k world (for the ncu capture of the DMMA trailing update)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs
NP, NL, E = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (10000, 2000, 100000)))
w = capi.synth_world(NP, NL, E, seed=0xB0500003)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"], fixed_pose_id=int(w["pose_ids"][0]))
ctx = capi.Context(solver=capi.SOLVER_DENSE_CHOLESKY)
pr.upload(ctx)
ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
ctx.triangulate()
for _ in range(2):
    s = ctx.step()
    print("dense n=%d: solve %.1f ms (%.2f TFLOP/s on n^3/3), launches %d, status %d" % (3 * NP, s.ms_solve, (3 * NP) ** 3 / 3 / (s.ms_solve * 1e-3) / 1e12, s.gpu_launches, s.solver_status))
