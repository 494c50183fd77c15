"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN CODE (oracle/_ref/libbos_ref.so: the reference's unmodified
sources compiled against oracle/eigen_standin, see oracle/Makefile target `ref`).  Build container only (needs /root/reference):

    python tests/golden/make_ref_golden.py

Cases: the two bundled datasets and their ground-truth flavours (parsed by the reference's own parse_g2o from /root/reference/data; the
GT files carry VERTEX_XY lines: landmarks in file order, no triangulation) and two small random worlds fed
through State::add_pose / the observation constructors (shuffled non-contiguous ids, a fixed pose in mid-chain, per-edge bearing
omegas, full 3x3 odometry omegas, a loop closure with source > destination, outliers that trigger the threshold kernel, a landmark
seen once).  Stored per case: the problem (random worlds only; mini / full are in mini.npz / full.npz), the id tables and the
triangulated landmarks, per-edge errors and Jacobians (analytic and numeric) at iteration 0, H_nofixed (scalar CSC) and b_nofixed as
left by the first step(), chi2 before every step and the state after the steps listed in `checkpoints`.  All values are the reference's floats.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

DATA = os.path.join(ref.REFERENCE_ROOT, "data")
OUT = os.path.dirname(os.path.abspath(__file__))
G2O = {"mini": ("mini_initial_guess.g2o", 50), "full": ("slam2D_bearing_only_initial_guess.g2o", 20),
       # the ground-truth files carry VERTEX_XY lines: landmarks enter the state in FILE order (no triangulation), the start is the GT state
       "mini_gt": ("mini_ground_truth.g2o", 10), "full_gt": ("slam2D_bearing_only_ground_truth.g2o", 10)}
CHECKPOINTS = [1, 2, 5, 10, 20, 50]


def random_world(seed, n_poses, n_lms, edges_per_pose, outliers, kernel_threshold=1.0):
    """A small world in float32-exact numbers: a noisy loop trajectory, landmarks around it, bearings with noise and a few gross outliers."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 2 * np.pi, n_poses, endpoint=False)
    true = np.stack([6 * np.cos(t), 4 * np.sin(t), t + np.pi / 2 + 0.1 * rng.standard_normal(n_poses)], 1)
    lms = np.stack([9 * rng.uniform(-1, 1, n_lms), 7 * rng.uniform(-1, 1, n_lms)], 1)
    pose_ids = rng.permutation(np.arange(1000, 1000 + 3 * n_poses, 3))[:n_poses].astype(np.int32)
    lm_ids = rng.permutation(np.arange(5, 5 + 7 * n_lms, 7))[:n_lms].astype(np.int32)

    def rel(a, b):   # odometry as the reference predicts it: R_a^T (t_b - t_a), theta_b - theta_a
        c, s = np.cos(a[2]), np.sin(a[2])
        d = b[:2] - a[:2]
        return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], (b[2] - a[2] + np.pi) % (2 * np.pi) - np.pi])

    o_src, o_dst, o_z, o_om = [], [], [], []
    for i in range(n_poses - 1):
        o_src.append(i); o_dst.append(i + 1)
        o_z.append(rel(true[i], true[i + 1]) + np.array([0.05, 0.05, 0.02]) * rng.standard_normal(3))
    o_src.append(n_poses - 1); o_dst.append(0)                      # loop closure, source stix > destination stix
    o_z.append(rel(true[-1], true[0]) + np.array([0.05, 0.05, 0.02]) * rng.standard_normal(3))
    for _ in o_src:
        a = rng.standard_normal((3, 3))
        o_om.append((a @ a.T + np.diag([50.0, 50.0, 200.0])).reshape(9))   # SPD with off-diagonal terms
    init = np.zeros_like(true)                                       # dead reckoning from the first pose
    init[0] = true[0]
    for i in range(n_poses - 1):
        c, s = np.cos(init[i, 2]), np.sin(init[i, 2])
        z = o_z[i]
        init[i + 1] = [init[i, 0] + c * z[0] - s * z[1], init[i, 1] + s * z[0] + c * z[1], init[i, 2] + z[2]]
    b_pose, b_lm, b_z, b_om = [], [], [], []
    for i in range(n_poses):
        seen = rng.choice(n_lms - 1, size=min(edges_per_pose, n_lms - 1), replace=False)
        for l in seen:
            d = lms[l] - true[i, :2]
            b_pose.append(i); b_lm.append(l)
            b_z.append(np.arctan2(d[1], d[0]) - true[i, 2] + 0.01 * rng.standard_normal())
            b_om.append(rng.uniform(0.5, 4.0))
    b_pose.append(n_poses // 3); b_lm.append(n_lms - 1)              # the last landmark is seen exactly once
    d = lms[-1] - true[n_poses // 3, :2]
    b_z.append(np.arctan2(d[1], d[0]) - true[n_poses // 3, 2]); b_om.append(1.0)
    b_z = np.array(b_z)
    bad = rng.choice(len(b_z) - 1, size=outliers, replace=False)
    b_z[bad] += rng.uniform(1.0, 2.5, outliers) * rng.choice([-1, 1], outliers)   # over the kernel threshold
    order = rng.permutation(len(b_z))                                # observations arrive in no particular order
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    return dict(pose_ids=pose_ids, poses_xyt=f32(init), lm_ids_true=lm_ids, lms_true=f32(lms),
                b_pose_id=pose_ids[np.array(b_pose)[order]], b_lm_id=lm_ids[np.array(b_lm)[order]], b_z=f32(b_z[order]),
                b_omega=f32(np.array(b_om)[order]), o_src_id=pose_ids[o_src], o_dst_id=pose_ids[o_dst], o_z=f32(o_z), o_omega=f32(o_om),
                fixed_pose_id=np.int32(pose_ids[n_poses // 2]), kernel_threshold=np.float32(kernel_threshold), damping=np.float32(0.01))


RANDOM = {"rand_a": dict(seed=11, n_poses=24, n_lms=9, edges_per_pose=4, outliers=6),
          "rand_b": dict(seed=23, n_poses=60, n_lms=25, edges_per_pose=6, outliers=20, kernel_threshold=0.5)}


def chi2(eb, eo, b_omega, o_omega):
    we = np.einsum("eij,ej->ei", o_omega.reshape(-1, 3, 3).astype(np.float64), eo.astype(np.float64))
    return float((b_omega * eb.astype(np.float64) ** 2).sum()), float(np.einsum("ei,ei->", we, eo.astype(np.float64)))


def run(name):
    r = ref.Reference()
    out = {}
    if name in G2O:
        fname, iters = G2O[name]
        r.load_g2o(os.path.join(DATA, fname))
        w = dict(np.load(os.path.join(OUT, name.replace("_gt", "") + ".npz")))
        b_omega = np.ones(len(w["b_z"])); o_omega = w["o_omega"]
        out["bound"] = np.float32(r.bound())
    else:
        w = random_world(**RANDOM[name])
        iters = 20
        r.set_problem(w["pose_ids"], w["poses_xyt"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"],
                      w["o_omega"], b_omega=w["b_omega"], fixed_id=int(w["fixed_pose_id"]))
        b_omega, o_omega = w["b_omega"], w["o_omega"]
        out.update({"problem_" + k: v for k, v in w.items()})
    if not name.endswith("_gt"):
        r.triangulate()
    pid, lid = r.ids()
    out["pose_ids"], out["lm_ids"] = pid, lid
    r.solver_init(-1)
    if name not in G2O:
        r.set_params(float(w["kernel_threshold"]), float(w["damping"]))
    out["fixed_pose_id"] = np.int32(r.counts()["fixed_pose_id"])
    out["poses0_xycs"], out["lms_tri"] = r.state()
    out["eb"], out["jb"], out["eo"], out["jo"] = r.edge_terms()
    _, out["jb_num"], _, out["jo_num"] = r.edge_terms(numeric=True)
    chis, rcs = [], []
    for it in range(1, iters + 1):
        eb, _, eo, _ = r.edge_terms()
        chis.append(chi2(eb, eo, b_omega, o_omega))
        rcs.append(r.step())
        if it == 1:
            out["H_colptr"], out["H_rowidx"], out["H_val"] = r.H(True)
            out["b_nofixed"] = r.b(True)
            out["H_full_nnz"] = np.int64(len(r.H(False)[1]))
        if it in CHECKPOINTS:
            out["P_it%d" % it], out["L_it%d" % it] = r.state()
    out["chi2"] = np.array(chis)
    out["step_rc"] = np.array(rcs, np.int32)
    out["checkpoints"] = np.array([c for c in CHECKPOINTS if c <= iters], np.int32)
    np.savez_compressed(os.path.join(OUT, "ref_" + name + ".npz"), **out)
    print(name, r.counts(), "chi2 first/last", chis[0], chis[-1], "not-SPD warnings", int(np.sum(np.array(rcs) == 2)))


if __name__ == "__main__":
    for n in sys.argv[1:] or list(G2O) + list(RANDOM):
        run(n)
