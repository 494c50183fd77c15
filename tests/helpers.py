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
    """column-norm-relative error of two CSC value arrays on the same pattern (vectorised: 30 M entries at synth-2M)"""
    colptr = np.asarray(colptr, np.int64)
    if len(val_b) == 0:
        return 0.0
    starts = colptr[:-1]
    nonempty = colptr[1:] > starts
    idx = starts[nonempty]
    den = np.maximum(np.maximum.reduceat(np.abs(val_b), idx), 1e-300)
    num = np.maximum.reduceat(np.abs(np.asarray(val_a) - np.asarray(val_b)), idx)
    return float((num / den).max())


def chi2_odometry_tolerance(pr, o, rel=1e-9, dtheta=4e-15):
    """Absolute tolerance for chi2_odometry = sum e^T Omega e.  On a dead-reckoned initial guess the odometry residuals are
    float-rounding residues (|e_theta| ~ 1e-6), so the few-ulp disagreement between CUDA's and glibc's atan2 on the pose angles
    (up to 4e-15 rad per residual: two angles of magnitude <= pi, 2 ulp each) is visible at ~1e-9 RELATIVE in the sum although
    every term is as accurate as FP64 allows.  Tolerance = rel * chi2 + sum_e 2 |(Omega e)_theta| * dtheta, from the oracle's terms."""
    eo = o.edge_terms()[2]
    we = np.einsum("eij,ej->ei", pr.o_omega.reshape(-1, 3, 3), eo)
    return rel * float(np.einsum("ei,ei->", we, eo)) + float((2 * np.abs(we[:, 2])).sum()) * dtheta


def align_oracle_wrap_branch(o, eb_gpu):
    """The +-pi policy: a bearing residual within 1e-9 of +-pi may wrap either way depending on the last bit
    of atan2 -- in the reference's FP32 arithmetic as much as in FP64.  Puts the oracle on the branch the device took for exactly
    those edges (test hook Oracle.set_wrap_branch) and returns their indices; everything else is compared unmodified."""
    eb_orc = o.edge_terms()[0]
    amb = np.where(np.abs(np.abs(eb_orc) - np.pi) < 1e-9)[0]
    assert np.all(np.abs(np.abs(eb_gpu[amb]) - np.pi) < 1e-9)
    o.set_wrap_branch(amb, np.where(eb_gpu[amb] >= 0, 1, -1))
    return amb


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
