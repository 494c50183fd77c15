"""torchrun --nproc-per-node N tools/diag_mgpu_steps.py : per-step PCG iterations / status of the sharded build + replicated solve."""
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
w = capi.synth_world(200000, 50000, 2000000, seed=0xB0500000)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
             fixed_pose_id=int(w["pose_ids"][0]))
for precond, mode in ((0, 2), (0, 1), (2, 2)):
    ctx = capi.Context(device=local, solver=capi.SOLVER_PCG, pcg_rtol=1e-8, pcg_max_iters=20000, pcg_precond=precond)
    pr.upload(ctx)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    ctx.set_reduce_mode(mode)
    ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
    ctx.triangulate()
    out = []
    for _ in range(7):
        s = ctx.step()
        out.append((s.pcg_iterations, s.solver_status, round(s.chi2_bearing + s.chi2_odometry, 1)))
    print("rank", rank, "precond", precond, "mode", mode, out, flush=True)
    ctx.close()
dist.destroy_process_group()
