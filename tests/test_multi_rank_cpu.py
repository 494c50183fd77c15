"""Multi-rank host logic on CPU (gloo, world_size 2): the edge-shard partition of include/bos_b200.h (bos_host_edge_shard)
and the "partials + allreduce = full H, b" property the multi-GPU path relies on.  Every rank assembles the
block-sparse H, b of ITS edge shard from the oracle's per-edge errors / Jacobians with numpy (damping on rank 0 only, as
the CUDA path does), the partials are summed with a gloo allreduce, and every rank must end up with the oracle's full blocks.
The bearing shard is over the (pose, landmark)-SORTED edge order in whole tiles, exactly what the GPU ranks linearize."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import golden_problem, load_golden, oracle_for
from prb_project_bearing_only_slam_b200 import capi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_blocks(pr, o, ob, rank, world, damping, kernel_threshold=1.0):
    """numpy assembly of one rank's shard: diagonal blocks Hpp 3x3 / Hll 2x2, b, and every off-diagonal block H[lo][hi] =
    J_lo^T Omega J_hi (pose-pose and pose-landmark; nodes = poses then landmarks) in the oracle's sorted block list.  Like the
    oracle (and the reference, slam/solver.cpp:72-73) the fixed pose's rows / columns are still present here: the gauge is
    applied afterwards by deleting them."""
    eb, jb, eo, jo = o.edge_terms()
    NP, NL = pr.NP, pr.NL
    Hpp = np.zeros((NP, 3, 3)); Hll = np.zeros((NL, 2, 2)); b = np.zeros(3 * NP + 2 * NL)
    key = {(int(a), int(c)): k for k, (a, c) in enumerate(zip(ob["off_lo"], ob["off_hi"]))}
    Hoff = np.zeros((len(key), 9))
    if rank == 0:
        Hpp += damping * np.eye(3); Hll += damping * np.eye(2)
    b0, b1, o0, o1 = capi.host_edge_shard(pr.Eb, pr.Eo, rank, world)
    order = np.lexsort((pr.b_lm, pr.b_pose))                       # the product's sorted edge order (stable)
    for e in order[b0:b1]:
        p, l = pr.b_pose[e], pr.b_lm[e]
        om = 1.0 if pr.b_omega is None else pr.b_omega[e]
        J = jb[e].copy(); err = eb[e]
        chi = err * om * err
        if chi > kernel_threshold:
            err *= np.sqrt(kernel_threshold / chi)
        Hpp[p] += om * np.outer(J[:3], J[:3]); Hll[l] += om * np.outer(J[3:], J[3:])
        b[3 * p:3 * p + 3] += J[:3] * om * err; b[3 * NP + 2 * l:3 * NP + 2 * l + 2] += J[3:] * om * err
        blk = np.zeros(9); blk[:6] = (om * np.outer(J[:3], J[3:])).ravel()
        Hoff[key[(p, NP + l)]] += blk
    for e in range(o0, o1):
        s, t = pr.o_src[e], pr.o_dst[e]
        Om = pr.o_omega[e].reshape(3, 3); J = jo[e].reshape(3, 6).copy(); err = eo[e].copy()
        chi = err @ Om @ err
        if chi > kernel_threshold:
            err *= np.sqrt(kernel_threshold / chi)
        Js, Jt = J[:, :3], J[:, 3:]
        Hpp[s] += Js.T @ Om @ Js; Hpp[t] += Jt.T @ Om @ Jt
        b[3 * s:3 * s + 3] += Js.T @ Om @ err; b[3 * t:3 * t + 3] += Jt.T @ Om @ err
        lo, hi = min(s, t), max(s, t)
        Hoff[key[(lo, hi)]] += ((Js.T @ Om @ Jt) if s == lo else (Jt.T @ Om @ Js)).ravel()
    return Hpp, Hll, Hoff, b


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        pr = golden_problem(g)
        o = oracle_for(g["pose_ids"], g["poses_xyt"], pr)
        o.step(0); o.step(0)                                        # away from the +-pi branch cut of the triangulated start
        o.linearize()
        ob = o.blocks()
        parts = _partial_blocks(pr, o, ob, rank, world, damping=float(np.float32(0.01)))
        tens = [torch.from_numpy(np.ascontiguousarray(x)) for x in parts]
        for t in tens:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        Hpp, Hll, Hoff, b = [t.numpy() for t in tens]
        den = lambda x: max(np.abs(x).max(), 1e-300)
        errs = [np.abs(Hpp.reshape(-1, 9) - ob["Hpp"]).max() / den(ob["Hpp"]), np.abs(Hll.reshape(-1, 4) - ob["Hll"]).max() / den(ob["Hll"]),
                np.abs(Hoff.reshape(-1, 9) - ob["Hoff"]).max() / den(ob["Hoff"]), np.abs(b - ob["b"]).max() / den(ob["b"])]
        shard = capi.host_edge_shard(pr.Eb, pr.Eo, rank, world)
        out.put((rank, errs, shard))
    except Exception:
        import traceback
        out.put((rank, traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["mini", "full"])
def test_sharded_partials_allreduce_to_the_full_system(built_lib, name):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    for rank, errs, shard in res:
        assert shard is not None, errs
        assert max(errs) <= 1e-12, (rank, errs)
    # the two shards tile the edges; bearing shards are whole 512-edge tiles of the sorted order
    (b0, b1, o0, o1), (c0, c1, p0, p1) = res[0][2], res[1][2]
    g = load_golden(name)
    assert b0 == 0 and b1 == c0 and c1 == len(g["b_z"]) and o0 == 0 and o1 == p0 and p1 == len(g["o_src_id"])
    assert b1 % 512 == 0 or b1 == len(g["b_z"])
