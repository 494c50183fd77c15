"""Generates tests/golden/*.npz from the reference's bundled datasets (read-only, /root/reference/data)
through the CPU oracle.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the parsed problem (values exactly as std::stof reads them), and the oracle's
FP64 and FP32 results: triangulated landmarks, iteration-0 H (scalar CSC of H_nofixed), b, per-edge
errors/Jacobians, per-iteration chi2 / |dx|, and the state after the named iteration count.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"mini": ("mini_initial_guess.g2o", "mini_ground_truth.g2o", 50),
         "full": ("slam2D_bearing_only_initial_guess.g2o", "slam2D_bearing_only_ground_truth.g2o", 30)}


def run(name, ig, gt, iters):
    out = {}
    for dt in ("f64", "f32"):
        o = Oracle(dt)
        o.load_g2o(os.path.join(REF, ig))
        c = o.counts()
        if dt == "f64":
            pid, _ = o.ids()
            e = o.edges()
            out.update(pose_ids=pid, poses_xyt=o.state_xyt(), fixed_pose_id=np.int32(c["fixed_pose_id"]), bound=np.float32(o.bound()),
                       b_pose_id=e["b_pose_id"], b_lm_id=e["b_lm_id"], b_z=e["b_z"], o_src_id=e["o_src_id"], o_dst_id=e["o_dst_id"],
                       o_z=e["o_z"], o_omega=e["o_omega"], poses_xycs=o.state()[0])
        o.triangulate()
        _, lid = o.ids()
        P0, L0 = o.state()
        out["lm_ids"] = lid
        out["single_obs_%s" % dt] = np.array(o.single_observation_landmarks(), np.int32)
        out["lms_tri_%s" % dt] = L0
        o.solver_init(c["fixed_pose_id"])
        o.linearize()
        colptr, rowidx, val, b = o.csc()
        eb, jb, eo, jo = o.edge_terms()
        out["csc_colptr"] = colptr; out["csc_rowidx"] = rowidx
        out["csc_val_%s" % dt] = val; out["b_nofixed_%s" % dt] = b
        out["err_b_%s" % dt] = eb; out["jac_b_%s" % dt] = jb; out["err_o_%s" % dt] = eo; out["jac_o_%s" % dt] = jo
        chi = []
        for it in range(iters):
            o.step(0)
            s = o.stats()
            chi.append([s["chi2_bearing"], s["chi2_odometry"], s["over_bearing"], s["over_odometry"], s["delta_inf"]])
            if it == 0:
                out["delta0_%s" % dt] = o.delta()
        out["trajectory_%s" % dt] = np.array(chi)
        Pf, Lf = o.state()
        out["poses_final_%s" % dt] = Pf; out["lms_final_%s" % dt] = Lf
    out["iterations"] = np.int32(iters)
    # ground-truth file: plausibility fixture only (the optimum is not the GT state)
    g = Oracle("f64")
    g.load_g2o(os.path.join(REF, gt))
    gp, gl = g.ids()
    out["gt_pose_ids"] = gp; out["gt_poses_xyt"] = g.state_xyt(); out["gt_lm_ids"] = gl; out["gt_lms_xy"] = g.state()[1]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith(("csc", "traj"))})


if __name__ == "__main__":
    for n, (ig, gt, it) in CASES.items():
        run(n, ig, gt, it)
