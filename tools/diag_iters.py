import os, sys
sys.path.insert(0, "/root/repo")
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs
w = capi.synth_world(200000, 50000, 2000000, seed=0xB0500000)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"], fixed_pose_id=int(w["pose_ids"][0]))
for precond in (0, 2, 1):
    ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=1e-8, pcg_max_iters=20000, pcg_precond=precond)
    pr.upload(ctx)
    ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
    ctx.triangulate()
    out = []
    for _ in range(8 if precond != 1 else 3):
        s = ctx.step(); out.append((s.pcg_iterations, round(s.ms_solve, 2), round(s.chi2_bearing + s.chi2_odometry, 1)))
    print("precond", precond, out, flush=True)
