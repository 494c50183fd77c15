"""torchrun tools/nccl_diag.py : what transport NCCL uses on this box and how long an allreduce of the H, b prefix takes."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (1 << 10, 1 << 20, 4 << 20, 16 << 20):
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1) / 10
        print("allreduce %8.2f MB f64: %.3f ms  (bus %.1f GB/s)" % (n * 8 / 1e6, ms, 2 * (world - 1) / world * n * 8 / ms / 1e6), flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m | head -12")
dist.destroy_process_group()
