"""Bit-exact checks of the host-side integer work (no GPU): sparsity pattern, edge-to-block indexing,
landmark association, gauge removal, edge sharding -- product pattern builder vs the oracle."""
import numpy as np
import pytest

from helpers import load_golden, golden_problem, synth_problem
from oracle.oracle import Oracle
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem


def _oracle_pattern(pr, pose_ids, poses_xyt, lms=None):
    o = Oracle("f64")
    o.set_problem(pose_ids, poses_xyt, pr.pose_ids[pr.b_pose], pr.lm_ids[pr.b_lm], pr.b_z, pr.pose_ids[pr.o_src],
                  pr.pose_ids[pr.o_dst], pr.o_z, pr.o_omega, fixed_id=pr.fixed_pose_id,
                  lm_ids=pr.lm_ids, lms_xy=np.zeros((pr.NL, 2)) if lms is None else lms)
    o.solver_init(pr.fixed_pose_id)
    o.linearize()
    return o


def _check(pr, pose_ids, poses_xyt):
    o = _oracle_pattern(pr, pose_ids, poses_xyt, lms=np.ones((pr.NL, 2)))
    hp = capi.HostPattern(pr.NP, pr.NL, pr.fixed_stix, pr.b_pose, pr.b_lm, pr.o_src, pr.o_dst)
    got = hp.get()
    colptr, rowidx, _, _ = o.csc()
    assert np.array_equal(got["csc_colptr"], colptr)
    assert np.array_equal(got["csc_rowidx"], rowidx)
    ob = o.blocks()
    # sorted set of block coordinates in the unified index space (pose i -> i, landmark j -> NP + j)
    mine = sorted([(int(a), int(b)) for a, b in zip(got["off_lo"], got["off_hi"])] +
                  [(int(a), pr.NP + int(b)) for a, b in zip(got["hpl_pose"], got["hpl_lm"])])
    theirs = [(int(a), int(b)) for a, b in zip(ob["off_lo"], ob["off_hi"])]
    assert mine == theirs
    # per-edge block: the slot of every edge names the same block coordinates as the oracle's edge -> stix resolution
    bp, bl, os_, od = o.edge_stix()
    assert np.array_equal(got["hpl_pose"][got["b_slot"]], bp) and np.array_equal(got["hpl_lm"][got["b_slot"]], bl)
    assert np.array_equal(got["off_lo"][got["o_slot"]], np.minimum(os_, od))
    assert np.array_equal(got["off_hi"][got["o_slot"]], np.maximum(os_, od))
    return got


@pytest.mark.parametrize("name", ["mini", "full"])
def test_pattern_bit_exact_on_bundled_data(built_lib, name):
    g = load_golden(name)
    pr = golden_problem(g)
    got = _check(pr, g["pose_ids"], g["poses_xyt"])
    assert np.array_equal(got["csc_colptr"], g["csc_colptr"]) and np.array_equal(got["csc_rowidx"], g["csc_rowidx"])
    # landmark association: stix = ascending id after triangulation (slam/triangulation.cpp:65-74)
    assert np.array_equal(pr.lm_ids, g["lm_ids"]) and np.all(np.diff(pr.lm_ids) > 0)


def test_pattern_bit_exact_on_synthetic_world(built_lib):
    w, pr = synth_problem(400, 90, 4000, seed=5)
    _check(pr, w["pose_ids"], w["poses_init"])


def test_pattern_with_duplicates_unsorted_edges_loops_and_isolated_blocks(built_lib):
    rng = np.random.default_rng(3)
    NP, NL = 12, 7
    pose_ids = np.array([897, 357, 205] + list(range(10, 19)), np.int32)  # non-contiguous ids (tests/state_test.cpp)
    lm_ids = np.array([3, 35, 36, 90, 91, 200, 777], np.int32)
    bp = rng.integers(0, NP - 1, 40); bl = rng.integers(0, NL - 1, 40)   # pose NP-1 and landmark NL-1 stay edge-free
    bp[5], bl[5] = bp[4], bl[4]                                          # duplicate (pose, lm) pair
    src = np.array([0, 1, 2, 3, 4, 5, 9, 3, 1]); dst = np.array([1, 2, 3, 4, 5, 6, 2, 4, 0])  # closure, duplicate pair, reversed pair
    pr = Problem(pose_ids, pose_ids[bp], lm_ids[bl], rng.normal(size=40), pose_ids[src], pose_ids[dst],
                 rng.normal(size=(9, 3)), np.tile(np.diag([500.0, 500, 5000]).ravel(), (9, 1)), fixed_pose_id=205, lm_ids=lm_ids)
    assert pr.fixed_stix == 2
    got = _check(pr, pose_ids, rng.normal(size=(NP, 3)))
    assert len(got["hpl_pose"]) < 40  # the duplicate collapsed into one block
    with pytest.raises(KeyError):
        Problem(pose_ids, [4242], [3], [0.1], [], [], np.zeros((0, 3)), np.zeros((0, 9)))


def test_invalid_problems_are_rejected(built_lib):
    with pytest.raises(capi.BosError):
        capi.HostPattern(3, 2, 0, [0, 3], [0, 1], [], [])          # pose index out of range
    with pytest.raises(capi.BosError):
        capi.HostPattern(3, 2, 0, [0], [0], [1], [1])              # odometry self-loop
    with pytest.raises(capi.BosError):
        capi.HostPattern(3, 2, 5, [0], [0], [0], [1])              # fixed pose out of range
    hp = capi.HostPattern(3, 2, 1, [], [], [], [])                 # empty edge lists are legal: H is the damping only
    assert hp.info().csc_nnz == 3 * 2 + 2 * 2


def test_edge_shards_partition_the_edges(built_lib):
    for Eb, Eo, R in ((15, 2, 2), (2132, 300, 8), (1000003, 7, 4), (5, 0, 8)):
        pieces = [capi.host_edge_shard(Eb, Eo, r, R) for r in range(R)]
        assert pieces[0][0] == 0 and pieces[-1][1] == Eb and pieces[0][2] == 0 and pieces[-1][3] == Eo
        for a, b in zip(pieces[:-1], pieces[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        sizes = [p[1] - p[0] for p in pieces]
        assert max(sizes) - min(s for s in sizes if s or True) <= max(sizes)  # contiguous, equal chunks except the tail
        assert len({s for s in sizes[:-1] if s}) <= 2


def test_threaded_pattern_build_equals_the_serial_one(built_lib, monkeypatch):
    """The tile grouping and the odometry tables are built on their own threads beside the ELL / chunk layouts; every integer
    table the device receives must be identical to a build that stays on one thread (BOS_PATTERN_THREADS=1)."""
    rng = np.random.default_rng(21)
    worlds = [synth_problem(4000, 900, 40000, seed=8)[1], synth_problem(333, 71, 2500, seed=9)[1]]
    for pr in worlds:
        # shuffled caller order and a few duplicate edges exercise the sort path as well
        perm = rng.permutation(len(pr.b_pose))
        bp = np.concatenate([pr.b_pose[perm], pr.b_pose[:17]]); bl = np.concatenate([pr.b_lm[perm], pr.b_lm[:17]])
        sums = []
        for threads in ("1", "8"):
            monkeypatch.setenv("BOS_PATTERN_THREADS", threads)
            hp = capi.HostPattern(pr.NP, pr.NL, pr.fixed_stix, bp, bl, pr.o_src, pr.o_dst)
            sums.append(hp.checksum())
        assert sums[0] == sums[1] and sums[0] != 0


def _schur_pose_pattern(pr):
    """Pose pairs of the reduced (pose-only Schur) system: poses sharing a landmark or an odometry edge (dense boolean, small worlds)."""
    A = np.zeros((pr.NP, pr.NP), bool)
    for l in range(pr.NL):
        ps = np.unique(pr.b_pose[pr.b_lm == l])
        A[np.ix_(ps, ps)] = True
    A[pr.o_src, pr.o_dst] = True
    A[pr.o_dst, pr.o_src] = True
    A[np.arange(pr.NP), np.arange(pr.NP)] = True
    return A


@pytest.mark.parametrize("case", ["full", "synth"])
def test_skyline_symbolic_phase_covers_the_factor(built_lib, case):
    """The symbolic phase of BOS_SOLVER_SPARSE_CHOLESKY (analyzePattern's analogue): its per-panel row limits are non-decreasing, contain every
    entry of the reduced system AND of its Cholesky factor (checked with a dense symbolic elimination), and every window the blocked
    factorisation touches fits into the stored rows per column."""
    if case == "full":
        g = load_golden("full")
        pr = golden_problem(g)
    else:
        _, pr = synth_problem(500, 120, 5000, seed=11)
    hp = capi.HostPattern(pr.NP, pr.NL, pr.fixed_stix, pr.b_pose, pr.b_lm, pr.o_src, pr.o_dst)
    pe, W, fill = hp.skyline()
    n = 3 * pr.NP
    assert len(pe) == (n + 63) // 64 and np.all(np.diff(pe) >= 0) and pe[-1] == n and 0 < fill <= 1.0 + 1e-12
    A = np.kron(_schur_pose_pattern(pr), np.ones((3, 3), bool))
    # symbolic Cholesky (no cancellation): L's pattern column by column
    Lp = np.tril(A)
    for j in range(n):
        rows = np.nonzero(Lp[j + 1:, j])[0] + j + 1
        if len(rows):
            Lp[np.ix_(rows, rows)] |= np.tril(np.ones((len(rows), len(rows)), bool))
    last = np.array([np.nonzero(Lp[:, j])[0].max() + 1 for j in range(n)])
    assert np.all(last <= pe[np.arange(n) // 64])
    for c0 in range(0, n, 256):
        cend = min(n, c0 + 256)
        assert pe[(cend - 1) // 64] - c0 <= W - 1 or W == n + 1
    if case == "synth":
        assert fill < 0.9          # a trajectory-ordered world has a real envelope
