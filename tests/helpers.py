"""Shared test plumbing: golden fixtures, synthetic problems, oracle construction."""
import os

import numpy as np

from oracle.oracle import Oracle
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_problem(g):
    return Problem(g["pose_ids"], g["b_pose_id"], g["b_lm_id"], g["b_z"], g["o_src_id"], g["o_dst_id"], g["o_z"], g["o_omega"],
                   fixed_pose_id=int(g["fixed_pose_id"]))


def synth_problem(n_poses, n_landmarks, edges, seed=1234, **kw):
    w = capi.synth_world(n_poses, n_landmarks, edges, seed=seed, **kw)
    pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
                 fixed_pose_id=int(w["pose_ids"][0]))
    return w, pr


def oracle_for(pose_ids, poses_xyt, pr, dtype="f64", b_pose_id=None, b_lm_id=None, o_src_id=None, o_dst_id=None,
               triangulate=True, lms=None):
    """Oracle loaded with the same problem.  Edge ids are reconstructed from the stix arrays."""
    o = Oracle(dtype)
    o.set_problem(pose_ids, poses_xyt, pr.pose_ids[pr.b_pose], pr.lm_ids[pr.b_lm], pr.b_z, pr.pose_ids[pr.o_src], pr.pose_ids[pr.o_dst],
                  pr.o_z, pr.o_omega, b_omega=pr.b_omega, fixed_id=pr.fixed_pose_id,
                  lm_ids=None if triangulate else pr.lm_ids, lms_xy=None if triangulate else lms)
    if triangulate:
        o.triangulate()
    o.solver_init(pr.fixed_pose_id)
    return o


def rel_block_err(a, b):
    """max over blocks of |a-b|_inf / max(|b|_inf, tiny): block-norm-relative error."""
    a = np.asarray(a); b = np.asarray(b)
    if a.size == 0:
        return 0.0
    a2 = a.reshape(a.shape[0], -1); b2 = b.reshape(b.shape[0], -1)
    den = np.maximum(np.abs(b2).max(axis=1), 1e-300)
    return float((np.abs(a2 - b2).max(axis=1) / den).max())


def csc_rel_err(colptr, val_a, val_b):
    """column-norm-relative error of two CSC value arrays on the same pattern"""
    worst = 0.0
    for j in range(len(colptr) - 1):
        s, e = colptr[j], colptr[j + 1]
        if e > s:
            den = max(np.abs(val_b[s:e]).max(), 1e-300)
            worst = max(worst, float(np.abs(val_a[s:e] - val_b[s:e]).max() / den))
    return worst


def angle_diff(a, b):
    d = np.asarray(a) - np.asarray(b)
    return (d + np.pi) % (2 * np.pi) - np.pi


def write_g2o_from_golden(g, path, ground_truth=False):
    """Writes a g2o file holding the golden fixture's problem (values are float32-exact, %.9g round-trips through std::stof):
    the initial-guess flavour (poses, FIX, edges) or the ground-truth flavour (landmarks + poses of the GT file, same edges)."""
    with open(path, "w") as f:
        if ground_truth:
            for i, (x, y) in zip(g["gt_lm_ids"], g["gt_lms_xy"]):
                f.write("VERTEX_XY %d %.9g %.9g\n" % (i, x, y))
            ids, xyt = g["gt_pose_ids"], g["gt_poses_xyt"]
        else:
            ids, xyt = g["pose_ids"], g["poses_xyt"]
        for i, (x, y, t) in zip(ids, xyt):
            f.write("VERTEX_SE2 %d %.9g %.9g %.9g\n" % (i, x, y, t))
        f.write("FIX %d\n" % int(g["fixed_pose_id"]))
        for s, d, z, om in zip(g["o_src_id"], g["o_dst_id"], g["o_z"], g["o_omega"].reshape(-1, 3, 3)):
            f.write("EDGE_SE2 %d %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n" % (s, d, z[0], z[1], z[2], om[0, 0], om[0, 1], om[0, 2],
                                                                                       om[1, 1], om[1, 2], om[2, 2]))
        for p, l, z in zip(g["b_pose_id"], g["b_lm_id"], g["b_z"]):
            f.write("EDGE_BEARING_SE2_XY %d %d %.9g 57295.8\n" % (p, l, z))
        f.write("\nSOMETHING_ELSE 1 2 3\n")     # the loader prints "Unrecognized SOMETHING_ELSE" and carries on
    return path
