"""torchrun --nproc-per-node N tools/mgpu_check.py : edge-sharded linearization + NCCL combine against the single-GPU result.
Every rank builds the same synthetic world; rank r linearizes its edge shard on GPU r; after the combine every rank must
hold the same H, b, chi2 as an unsharded context (1e-12 relative: only the summation order differs), for every reduce mode
(0 full allreduce, 1 allreduce of the overlapping blocks + allgather of the pose-landmark planes, 2 allreduce of the overlapping
blocks only, 3 ownership-based: landmark blocks summed, owned pose ranges gathered, 4 the same ownership with the bearing kernel writing
straight into every rank's replica over NVLink peer memory, no collective, 5 a local build followed by bulk pulls over the same peer
mappings), and a full step must give the same state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NP, NL, E = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (3000, 700, 30000)))
w = capi.synth_world(NP, NL, E, seed=99)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
             fixed_pose_id=int(w["pose_ids"][0]))
ref = capi.Context(device=local, solver=capi.SOLVER_PCG, pcg_rtol=1e-12, pcg_max_iters=20000)
pr.upload(ref)
ref.set_state(xyt_to_xycs(w["poses_init"]), None)
ref.triangulate()
P0, L0 = ref.get_state()
ref.linearize()
rb = ref.blocks(); rs = ref.stats()
ok = True
for mode in (0, 1, 2, 3, 4, 5):
    ctx = capi.Context(device=local, solver=capi.SOLVER_PCG, pcg_rtol=1e-12, pcg_max_iters=20000)
    pr.upload(ctx)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    if mode >= 4:                      # combine over NVLink peer memory (4: pushed by the bearing kernel, 5: pulled in bulk): exchange the CUDA IPC handles
        ctx.peer_connect(dist)
    ctx.set_reduce_mode(mode)
    ctx.set_state(P0, L0)
    ctx.linearize()
    b = ctx.blocks(); s = ctx.stats()
    for k in ("Hpp", "Hll", "Hoff", "b") + (("Hpl",) if mode < 2 else ()):   # modes 2 / 3 leave the pose-landmark blocks rank-local
        den = max(np.abs(rb[k]).max(), 1e-300)
        err = np.abs(b[k] - rb[k]).max() / den
        if err > 1e-12:
            ok = False
            print("rank %d mode %d %s mismatch %.3e" % (rank, mode, k, err), flush=True)
    if abs(s.chi2_bearing - rs.chi2_bearing) > 1e-10 * rs.chi2_bearing or abs(s.chi2_odometry - rs.chi2_odometry) > 1e-10 * max(rs.chi2_odometry, 1e-300) \
            or s.over_bearing != rs.over_bearing or s.over_odometry != rs.over_odometry:
        ok = False
        print("rank %d mode %d chi2 mismatch" % (rank, mode), s.chi2_bearing, rs.chi2_bearing, s.chi2_odometry, rs.chi2_odometry, flush=True)
    if mode >= 4:                      # several builds in a row: the barrier epochs and the re-initialisation of the replicas hold up
        for _ in range(3):
            ctx.linearize()
        b2 = ctx.blocks()
        for k in ("Hpp", "Hll", "Hoff", "b"):
            if np.abs(b2[k] - rb[k]).max() > 1e-12 * max(np.abs(rb[k]).max(), 1e-300):
                ok = False
                print("rank %d mode %d repeated build: %s mismatch" % (rank, mode, k), flush=True)
    ctx.set_state(P0, L0); ref.set_state(P0, L0)
    st = ctx.step(); ref.step()
    Pa, La = ctx.get_state(); Pb, Lb = ref.get_state()
    d = max(np.abs(Pa - Pb).max(), np.abs(La - Lb).max())
    if d > 1e-8:
        ok = False
        print("rank %d mode %d state mismatch %.3e" % (rank, mode, d), flush=True)
    print("rank %d mode %d shard %s ms lin %.3f allreduce %.3f solve %.3f pcg %d" % (rank, mode, ctx.edge_shard(), st.ms_linearize, st.ms_allreduce,
                                                                              st.ms_solve, st.pcg_iterations), flush=True)
    ctx.close()
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU_CHECK", "OK" if t.item() == 1.0 else "FAILED", "world", world, flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
