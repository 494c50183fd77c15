"""On-device problem setup against the host builder: bit-exact tables (checksum over every pattern table), identical id -> stix
resolution (framework/state.cpp:20-67 semantics), identical GN steps."""
import time

import numpy as np
import pytest

from helpers import golden_problem, load_golden, oracle_for, synth_problem
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem

pytestmark = pytest.mark.gpu


def _both(pr, chunks=None):
    host = capi.HostPattern(pr.NP, pr.NL, pr.fixed_stix, pr.b_pose, pr.b_lm, pr.o_src, pr.o_dst)
    ctx_h = capi.Context(); ctx_h.set_device_setup(False); pr.upload(ctx_h)
    ctx_d = capi.Context(); ctx_d.set_device_setup(True); pr.upload(ctx_d)
    return host, ctx_h, ctx_d


@pytest.mark.parametrize("case", ["mini", "full", "synth", "shuffled-duplicates"])
def test_device_built_tables_are_bit_identical(built_lib, case):
    if case in ("mini", "full"):
        g = load_golden(case); pr = golden_problem(g)
    elif case == "synth":
        _, pr = synth_problem(5000, 1200, 50000, seed=17)
    else:   # caller order shuffled, duplicated (pose, landmark) pairs, unobserved landmarks, edge-free poses
        _, pr = synth_problem(600, 150, 6000, seed=23)
        rng = np.random.default_rng(5)
        perm = rng.permutation(pr.Eb)
        dup = rng.integers(0, pr.Eb, 200)
        idx = np.concatenate([perm, dup])
        pr.b_pose, pr.b_lm, pr.b_z = pr.b_pose[idx].copy(), pr.b_lm[idx].copy(), pr.b_z[idx].copy()
        if pr.b_omega is not None:
            pr.b_omega = pr.b_omega[idx].copy()
        pr.Eb = len(idx)
    host, ctx_h, ctx_d = _both(pr)
    # the host-built context and the host-only builder agree (same code), the device-built context must agree with both
    assert ctx_h.pattern_checksum() == ctx_d.pattern_checksum()
    dev_ms, host_ms = ctx_d.last_setup_ms()
    assert dev_ms > 0.0
    ctx_h.close(); ctx_d.close()


def test_device_setup_gives_the_same_gn_steps(built_lib):
    w, pr = synth_problem(5000, 1200, 50000, seed=17)
    o = oracle_for(w["pose_ids"], w["poses_init"], pr)
    P, L = o.state()
    outs = []
    for dev in (False, True):
        ctx = capi.Context(solver=capi.SOLVER_PCG, pcg_rtol=1e-12)
        ctx.set_device_setup(dev)
        pr.upload(ctx); ctx.set_state(P, L)
        for _ in range(3):
            s = ctx.step()
        outs.append((s.chi2_bearing + s.chi2_odometry, ctx.get_state()))
        ctx.close()
    assert outs[0][0] == pytest.approx(outs[1][0], rel=1e-9)
    assert np.abs(outs[0][1][0] - outs[1][1][0]).max() <= 1e-8 and np.abs(outs[0][1][1] - outs[1][1][1]).max() <= 1e-8


def test_device_id_resolution_matches_the_host_maps(built_lib):
    rng = np.random.default_rng(9)
    NP = 3000
    pose_ids = (1200 + rng.permutation(NP) * 3).astype(np.int32)        # insertion order is NOT id order
    pose_ids[17] = pose_ids[5]                                          # duplicated id: the map keeps the LAST insertion (state.cpp:23)
    lm_pool = rng.choice(np.arange(-50, 100000, 7), 400, replace=False).astype(np.int32)   # sparse, negative ids included
    Eb, Eo = 20000, 2999
    usable = np.setdiff1d(np.arange(NP), [5])
    b_pose_id = pose_ids[rng.choice(usable, Eb)]
    b_lm_id = lm_pool[rng.integers(0, 300, Eb)]                         # 100 landmarks of the pool are never observed
    o_src_id = pose_ids[rng.choice(usable, Eo)]; o_dst_id = pose_ids[rng.choice(usable, Eo)]
    ref = Problem(pose_ids, b_pose_id, b_lm_id, np.zeros(Eb), o_src_id, o_dst_id, np.zeros((Eo, 3)), np.tile(np.eye(3).ravel(), (Eo, 1)),
                  fixed_pose_id=int(pose_ids[0]))
    bp, bl, os_, od, lm_ids = capi.device_resolve_ids(pose_ids, b_pose_id, b_lm_id, o_src_id, o_dst_id)
    assert np.array_equal(bp, ref.b_pose) and np.array_equal(bl, ref.b_lm) and np.array_equal(os_, ref.o_src) and np.array_equal(od, ref.o_dst)
    assert np.array_equal(lm_ids, ref.lm_ids)
    # the duplicated id resolves to its last insertion
    q = capi.device_resolve_ids(pose_ids, pose_ids[[5]], b_lm_id[:1], o_src_id[:0], o_dst_id[:0])
    assert q[0][0] == 17
    with pytest.raises(capi.BosError):      # std::map::at throws on an unknown id (state.cpp:43-49)
        capi.device_resolve_ids(pose_ids, np.array([7], np.int32), b_lm_id[:1], o_src_id[:0], o_dst_id[:0])


def test_device_setup_at_2m_edges_is_timed(built_lib):
    w, pr = synth_problem(200000, 50000, 2000000, seed=0xB0500003)
    t = {}
    for dev in (False, True):
        ctx = capi.Context(); ctx.set_device_setup(dev)
        t0 = time.perf_counter(); pr.upload(ctx); t[dev] = time.perf_counter() - t0
        ms = ctx.last_setup_ms()
        cs = ctx.pattern_checksum()
        if dev:
            assert cs == t["cs"]
            print("upload at %d edges: host builder %.1f ms, device core %.1f ms + host remainder %.1f ms (whole upload %.1f vs %.1f ms)" % (
                pr.Eb + pr.Eo, t["host_ms"], ms[0], ms[1], 1e3 * t[True], 1e3 * t[False]))
        else:
            t["cs"] = cs; t["host_ms"] = ms[1]
        ctx.close()
