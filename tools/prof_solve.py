"""GPU box: one linearize + solve of a synthetic world with a capped PCG iteration count (for ncu captures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs

NP, NL, E = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (200000, 50000, 2000000)))
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 50
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
precond = int(sys.argv[6]) if len(sys.argv) > 6 else 0
w = capi.synth_world(NP, NL, E, seed=0xB0500003)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
             fixed_pose_id=int(w["pose_ids"][0]))
ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=1e-8, pcg_max_iters=iters, pcg_variant=variant, pcg_precond=precond)
pr.upload(ctx)
ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
ctx.triangulate()
for _ in range(2):
    s = ctx.step()
    print("pcg", s.pcg_iterations, "ms lin %.3f solve %.3f (%.2f us/iter) upd %.3f" % (s.ms_linearize, s.ms_solve, 1e3 * s.ms_solve / max(s.pcg_iterations, 1), s.ms_update))
