"""GPU box: time the H, b build (bos_linearize phases of bos_step are CUDA-event timed) on a synthetic world."""
import os, sys, statistics
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs

NP, NL, E = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (200000, 50000, 2000000)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
prec = capi.PRECISION_F32 if (len(sys.argv) > 5 and sys.argv[5] == "f32") else capi.PRECISION_F64
w = capi.synth_world(NP, NL, E, seed=0xB0500003)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
             fixed_pose_id=int(w["pose_ids"][0]))
ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=1e-8, pcg_max_iters=2, precision=prec)
pr.upload(ctx)
ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
ctx.triangulate()
P0, L0 = ctx.get_state()
ms = []
for i in range(reps):
    ctx.set_state(P0, L0)
    s = ctx.step()
    ms.append(s.ms_linearize)
ms = ms[3:]
print("linearize ms: median %.4f min %.4f max %.4f  (%d edges, %s)" % (statistics.median(ms), min(ms), max(ms), pr.Eb + pr.Eo, "f32" if prec else "f64"))
