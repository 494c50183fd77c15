"""GPU timeline of a few GN steps at synth-2M through torch.profiler (Kineto / CUPTI sees every kernel, memset and memcpy of the
process, ours included): where the time between the kernels of a step goes.  Usage: python tools/timeline.py [steps] -> gpurun_out/timeline.txt"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prb_project_bearing_only_slam_b200 import capi  # noqa: E402
from prb_project_bearing_only_slam_b200.problem import Problem, xyt_to_xycs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w = capi.synth_world(200000, 50000, 2000000)
pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
             fixed_pose_id=int(w["pose_ids"][0]))
torch.cuda.init()
ctx = capi.Context(device=0, solver=capi.SOLVER_PCG, pcg_rtol=1e-8)
pr.upload(ctx)
ctx.set_state(xyt_to_xycs(w["poses_init"]), None)
ctx.triangulate()
for _ in range(10):
    ctx.step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        ctx.step()
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "timeline.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
lines = []
prev_end = None
for e in ev:
    gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
    lines.append("%12.1f  gap %8.1f  dur %9.1f  %s" % (e["ts"] - ev[0]["ts"], gap, e["dur"], e["name"][:70]))
    prev_end = e["ts"] + e["dur"]
open(os.path.join(ROOT, "gpurun_out", "timeline.txt"), "w").write("\n".join(lines) + "\n")
os.remove(out)
print("\n".join(lines[-45:]))
