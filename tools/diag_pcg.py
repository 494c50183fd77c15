"""Diagnostic (GPU box): PCG iteration counts / residuals of the synthetic worlds, checked with scipy on the downloaded CSC."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs
import scipy.sparse as sp

def run(NP, NL, E, rtol, seed=0xB0500003, check=True):
    w = capi.synth_world(NP, NL, E, seed=seed)
    pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
                 fixed_pose_id=int(w["pose_ids"][0]))
    ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=rtol, pcg_max_iters=20000)
    pr.upload(ctx)
    ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
    ctx.triangulate()
    for it in range(6):
        if check and it in (0, 3):
            ctx.linearize(); ctx.solve()
            colptr, rowidx, val, b = ctx.csc()
            d = ctx.delta()
            keep = np.ones(len(d), bool); keep[3*pr.fixed_stix:3*pr.fixed_stix+3] = False
            n = len(colptr) - 1
            H = sp.csc_matrix((val, rowidx, colptr), shape=(n, n))
            r = H @ d[keep] + b
            print("  it %d: |H dx + b|inf / |b|inf = %.3e   |dx|inf = %.3e" % (it, np.abs(r).max() / np.abs(b).max(), np.abs(d).max()), flush=True)
        s = ctx.step()
        print("NP %d it %d chi2 %.6e + %.6e over %d/%d dinf %.3e pcg %d status %d  ms lin %.3f solve %.3f upd %.3f" % (
            NP, it, s.chi2_bearing, s.chi2_odometry, s.over_bearing, s.over_odometry, s.delta_inf, s.pcg_iterations, s.solver_status,
            s.ms_linearize, s.ms_solve, s.ms_update), flush=True)

if __name__ == "__main__":
    run(20000, 4000, 200000, 1e-8)
    run(200000, 50000, 2000000, 1e-8)
    run(200000, 50000, 2000000, 1e-12, check=False)
