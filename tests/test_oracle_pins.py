"""Pins the CPU oracle against every known answer the reference offers for this path.

The reference has no assertions and cannot be compiled here (Eigen3/OpenCV absent), so these are its
soft pins: source-comment values, bundled data facts, the README's convergence statement,
plus an independent numpy restatement of one linearization + solve.
"""
import math
import os

import numpy as np
import pytest

from helpers import load_golden, golden_problem, oracle_for
from oracle import oracle as orc
from oracle.oracle import Oracle

REF_DATA = "/root/reference/data"
have_ref = os.path.isdir(REF_DATA)


def test_predict_bearing_known_answers():
    # tests/solver_stuff.cpp:25-38
    for dt, tol in (("f64", 1e-15), ("f32", 1e-6)):
        o = Oracle(dt)
        pi = math.pi
        assert abs(o.predict_bearing(0, 0, 0, 1, 0) - 0.0) <= tol
        assert abs(o.predict_bearing(0, 0, 0, 0, 1) - pi / 2) <= tol
        assert abs(abs(o.predict_bearing(0, 0, 0, -1, 0)) - pi) <= tol
        assert abs(o.predict_bearing(0, 0, 0, 0, -1) + pi / 2) <= tol
        assert abs(o.predict_bearing(0, 0, 0, 1, 1) - pi / 4) <= tol
        assert abs(o.predict_bearing(0, 0, pi / 2, 1, 1) + pi / 4) <= 10 * tol
        assert abs(abs(o.predict_bearing(0, 0, pi, 1, 0)) - pi) <= 10 * tol


def test_angle_helpers():
    pi = math.pi
    # normalized_angle: [-pi, pi)   (slam/solver_jacobians.cpp:325-333)
    assert orc.normalized_angle(pi) == pytest.approx(-pi)
    assert orc.normalized_angle(-pi) == pytest.approx(-pi)
    assert orc.normalized_angle(3 * pi + 0.25) == pytest.approx(-pi + 0.25)
    assert orc.normalized_angle(-7.0) == pytest.approx(-7.0 + 2 * pi)
    # smallestAngle: [-pi, pi] via fmod
    assert orc.smallest_angle(3.5) == pytest.approx(3.5 - 2 * pi)
    assert orc.smallest_angle(-3.5) == pytest.approx(-3.5 + 2 * pi)
    assert orc.smallest_angle(10.0) == pytest.approx(math.fmod(10.0, 2 * pi) - 2 * pi)
    assert orc.smallest_angle(1.0) == 1.0
    # float flavour narrows after every step
    assert orc.normalized_angle(4.0, "f32") == pytest.approx(4.0 - 2 * pi, abs=1e-6)


def test_colpiv_qr_matches_lstsq_and_rank1_rule():
    rng = np.random.default_rng(7)
    for M in (2, 3, 7, 60):
        A = rng.normal(size=(M, 2)); b = rng.normal(size=M)
        x = orc.colpiv_solve(A, b)
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-10, atol=1e-12)
    # M = 1: basic solution, pivot column gets b/a, the other coordinate stays 0 (slam/triangulation.cpp:38-42 case)
    x = orc.colpiv_solve(np.array([[0.3, -0.9]]), np.array([1.8]))
    assert x[0] == 0.0 and x[1] == pytest.approx(-2.0)
    x = orc.colpiv_solve(np.array([[0.9, -0.3]]), np.array([1.8]))
    assert x[1] == 0.0 and x[0] == pytest.approx(2.0)


@pytest.mark.parametrize("name", ["mini", "full"])
def test_golden_structure(name):
    g = load_golden(name)
    if name == "mini":
        assert (len(g["pose_ids"]), len(g["lm_ids"]), len(g["b_z"]), len(g["o_src_id"])) == (3, 6, 15, 2)
    else:
        assert (len(g["pose_ids"]), len(g["lm_ids"]), len(g["b_z"]), len(g["o_src_id"])) == (301, 141, 2132, 300)
        # slam/triangulation.cpp:41: landmarks 69, 112, 114 have a single observation
        assert g["single_obs_f64"].tolist() == [69, 112, 114] and g["single_obs_f32"].tolist() == [69, 112, 114]
        # scalar nnz of the full H is 34 257: H_nofixed + what the fixed pose's rows/cols held
        pr = golden_problem(g)
        k = int((pr.b_pose == pr.fixed_stix).sum())
        n_odo = int((pr.o_src == pr.fixed_stix).sum() + (pr.o_dst == pr.fixed_stix).sum())
        assert len(g["csc_rowidx"]) + 9 + 12 * k + 18 * n_odo == 34257
        assert pr.fixed_stix == 298
    assert int(g["fixed_pose_id"]) == 1498
    assert np.all(g["o_omega"][:, [0, 4, 8]] == [500, 500, 5000])


def test_convergence_matches_readme_and_survey_band():
    g = load_golden("full")
    t = g["trajectory_f64"]
    # SURVEY 6 sanity band: 96.8641 + 1.7e-4 at iteration 0 with 10 bearing edges over the threshold
    assert t[0, 0] == pytest.approx(96.8641, abs=2e-4) and t[0, 1] == pytest.approx(1.7e-4, abs=2e-5) and t[0, 2] == 10
    # README.md:22: "~20 iterations"; flat chi2 from there on
    assert abs(t[20, 0] - t[29, 0]) < 1e-3 and t[29, 0] == pytest.approx(4.105, abs=2e-3) and t[29, 1] == pytest.approx(1.778, abs=2e-3)
    m = load_golden("mini")["trajectory_f64"]
    assert m[0, 0] == pytest.approx(4.40679e-4, rel=1e-4) and m[-1, 0] == pytest.approx(4.2868e-4, rel=1e-4)
    # FP32 flavour follows the FP64 one at FP32 accuracy
    t32 = g["trajectory_f32"]
    assert t32[0, 0] == pytest.approx(t[0, 0], rel=1e-5) and t32[29, 0] == pytest.approx(t[29, 0], rel=2e-4)


@pytest.mark.skipif(not have_ref, reason="reference data only exists in the build container")
def test_jacobian_statistics_match_reference_comments():
    # tests/solver_stuff.cpp:82-88 (bearing, GT state) and :156-162 (odometry, IG + triangulated state), FP32, eps 1e-3
    o = Oracle("f32")
    o.load_g2o(os.path.join(REF_DATA, "slam2D_bearing_only_ground_truth.g2o"))
    c = o.counts()
    sums, maxs = [], []
    for e in range(c["Eb"]):
        a, n = o.bearing_jacobians(e)
        d = np.abs(a.astype(np.float32) - n.astype(np.float32))
        sums.append(d.sum()); maxs.append(d.max())
    assert max(sums) == pytest.approx(0.0135395, rel=0.15) and max(maxs) == pytest.approx(0.0131645, rel=0.15)
    assert np.mean(sums) == pytest.approx(0.000372358, rel=0.15) and np.mean(maxs) == pytest.approx(0.000166852, rel=0.15)
    o = Oracle("f32")
    o.load_g2o(os.path.join(REF_DATA, "slam2D_bearing_only_initial_guess.g2o"))
    o.triangulate()
    c = o.counts()
    sums, maxs = [], []
    for e in range(c["Eo"]):
        a, n = o.odometry_jacobians(e)
        d = np.abs(a.astype(np.float32) - n.astype(np.float32))
        sums.append(d.sum()); maxs.append(d.max())
    assert max(sums) == pytest.approx(0.00385106, rel=0.15) and max(maxs) == pytest.approx(0.000556946, rel=0.15)
    assert np.mean(sums) == pytest.approx(0.0017143, rel=0.15) and np.mean(maxs) == pytest.approx(0.000354741, rel=0.15)


@pytest.mark.skipif(not have_ref, reason="reference data only exists in the build container")
def test_predict_odometry_equals_measurement_on_initial_guess():
    # tests/solver_stuff.cpp:93-114: the initial guess is dead-reckoned from the odometry
    o = Oracle("f32")
    o.load_g2o(os.path.join(REF_DATA, "slam2D_bearing_only_initial_guess.g2o"))
    P = o.state_xyt(); pid, _ = o.ids(); e = o.edges()
    ix = {int(i): k for k, i in enumerate(pid)}
    for k in (0, 10, 42, 111, 128, 163, 222, 255):
        pred = o.predict_odometry(P[ix[int(e["o_src_id"][k])]], P[ix[int(e["o_dst_id"][k])]])
        d = pred - e["o_z"][k]
        d[2] = (d[2] + math.pi) % (2 * math.pi) - math.pi
        assert np.all(np.abs(d) < 2e-3)


@pytest.mark.skipif(not have_ref, reason="reference data only exists in the build container")
def test_golden_fixture_is_current():
    g = load_golden("mini")
    o = Oracle("f64")
    o.load_g2o(os.path.join(REF_DATA, "mini_initial_guess.g2o"))
    assert np.array_equal(o.state_xyt(), g["poses_xyt"]) and o.bound() == float(g["bound"]) == 12.0
    assert o.counts()["fixed_pose_id"] == 1498


def _numpy_step(g, lms):
    """Independent FP64 restatement (dense, numpy) of one linearization + solve: second opinion on the oracle."""
    pr = golden_problem(g)
    X = g["poses_xycs"]; NP, NL = pr.NP, pr.NL
    N = 3 * NP + 2 * NL
    H = np.zeros((N, N)); b = np.zeros(N); chi_b = chi_o = 0.0
    St = np.eye(N, dtype=bool)  # structural pattern: full blocks of every touched variable pair + the damping diagonal
    wrap = lambda a: (a + np.pi) % (2 * np.pi) - np.pi
    for e in range(pr.Eb):
        p, l = pr.b_pose[e], pr.b_lm[e]
        x, y, c, s = X[p]; R = np.array([[c, -s], [s, c]]); t = np.array([x, y]); lm = lms[l]
        gv = R.T @ (lm - t)
        err = wrap(math.atan2(gv[1], gv[0]) - g["b_z"][e])
        # a residual within rounding of +-pi can wrap either way (the bundled data has such edges: a two-observation
        # landmark triangulated BEHIND a pose); align the branch with the oracle, which follows the reference's
        # `while (angle >= CV_PI)` rule to the last bit.
        if abs(abs(err) - np.pi) < 1e-9:
            err = math.copysign(abs(err), g["err_b_f64"][e])
        a = np.array([-gv[1], gv[0]]) / (gv @ gv)
        J = np.zeros(N)
        J[3 * p:3 * p + 2] = a @ (-R.T); J[3 * p + 2] = a @ (R.T @ np.array([lm[1], -lm[0]]))
        J[3 * NP + 2 * l:3 * NP + 2 * l + 2] = a @ R.T
        chi = err * err; chi_b += chi
        if chi > 1.0:
            err *= math.sqrt(1.0 / chi)
        H += np.outer(J, J); b += J * err
        ix = np.r_[3 * p:3 * p + 3, 3 * NP + 2 * l:3 * NP + 2 * l + 2]; St[np.ix_(ix, ix)] = True
    for e in range(pr.Eo):
        s_, d_ = pr.o_src[e], pr.o_dst[e]
        xs, ys, cs, ss = X[s_]; xd, yd, cd, sd = X[d_]
        Rs = np.array([[cs, -ss], [ss, cs]]); td = np.array([xd, yd]); ts = np.array([xs, ys])
        pred = np.concatenate([Rs.T @ (td - ts), [wrap(math.atan2(sd, cd) - math.atan2(ss, cs))]])
        err = pred - g["o_z"][e]; err[2] = wrap(err[2])
        Om = g["o_omega"][e].reshape(3, 3)
        D = np.array([[0.0, -1.0], [1.0, 0.0]])
        J = np.zeros((3, N))
        J[:2, 3 * s_:3 * s_ + 2] = -Rs.T; J[:2, 3 * s_ + 2] = (D @ Rs).T @ td; J[2, 3 * s_ + 2] = -1
        J[:2, 3 * d_:3 * d_ + 2] = Rs.T; J[:2, 3 * d_ + 2] = Rs.T @ D @ td; J[2, 3 * d_ + 2] = 1
        chi = err @ Om @ err; chi_o += chi
        if chi > 1.0:
            err = err * math.sqrt(1.0 / chi)
        H += J.T @ Om @ J; b += J.T @ Om @ err
        ix = np.r_[3 * s_:3 * s_ + 3, 3 * d_:3 * d_ + 3]; St[np.ix_(ix, ix)] = True
    H += float(np.float32(0.01)) * np.eye(N)
    keep = np.ones(N, bool); keep[3 * pr.fixed_stix:3 * pr.fixed_stix + 3] = False
    dx = np.zeros(N); dx[keep] = np.linalg.solve(H[np.ix_(keep, keep)], -b[keep])
    return H[np.ix_(keep, keep)], b[keep], dx, chi_b, chi_o, St[np.ix_(keep, keep)]


@pytest.mark.parametrize("name", ["mini", "full"])
def test_oracle_agrees_with_independent_numpy_restatement(name):
    import scipy.sparse as sp
    g = load_golden(name)
    Hn, bn, dx, chi_b, chi_o, St = _numpy_step(g, g["lms_tri_f64"])
    n = len(g["csc_colptr"]) - 1
    Ho = sp.csc_matrix((g["csc_val_f64"], g["csc_rowidx"], g["csc_colptr"]), shape=(n, n)).toarray()
    So = sp.csc_matrix((np.ones(len(g["csc_rowidx"])), g["csc_rowidx"], g["csc_colptr"]), shape=(n, n)).toarray() != 0
    assert np.array_equal(So, St)  # sparsity pattern incl. explicit zeros: bit-exact
    for j in range(n):  # sorted row indices inside every column
        r = g["csc_rowidx"][g["csc_colptr"][j]:g["csc_colptr"][j + 1]]
        assert np.all(np.diff(r) > 0)
    assert np.abs(Ho - Hn).max() <= 1e-9 * np.abs(Hn).max()
    assert np.abs(g["b_nofixed_f64"] - bn).max() <= 1e-9 * max(np.abs(bn).max(), 1e-12)
    assert g["trajectory_f64"][0, 0] == pytest.approx(chi_b, rel=1e-10) and g["trajectory_f64"][0, 1] == pytest.approx(chi_o, rel=1e-8)
    assert np.abs(g["delta0_f64"] - dx).max() <= 1e-8 * np.abs(dx).max()


def test_irls_flavour_keeps_b_and_reweights_H():
    """The opt-in IRLS robust kernel (not in the reference): the threshold weight w = sqrt(kt / chi2) scales Omega instead of the error,
    so b is unchanged and H shrinks exactly on the blocks of the over-threshold edges."""
    g = load_golden("full")
    pr = golden_problem(g)
    o = oracle_for(g["pose_ids"], g["poses_xyt"], pr)
    for _ in range(2):       # the dead-reckoned start has zero odometry residuals
        o.step(0)
    o.set_params(1e-3, float(np.float32(0.01)))    # ... and at the reference's threshold of 1 no odometry edge of this trajectory is re-weighted
    o.linearize()
    col, row, v0, b0 = o.csc()
    o.set_irls(True)
    o.linearize()
    _, _, v1, b1 = o.csc()
    assert np.array_equal(b0, b1)
    assert o.stats()["over_odometry"] > 0 and o.stats()["over_bearing"] > 0
    diag0 = np.array([v0[col[j]:col[j + 1]][row[col[j]:col[j + 1]] == j][0] for j in range(len(col) - 1)])
    diag1 = np.array([v1[col[j]:col[j + 1]][row[col[j]:col[j + 1]] == j][0] for j in range(len(col) - 1)])
    assert np.all(diag1 <= diag0 + 1e-12) and np.any(diag1 < 0.9 * diag0)
    o.set_irls(False)
    o.linearize()
    assert np.array_equal(o.csc()[2], v0)
