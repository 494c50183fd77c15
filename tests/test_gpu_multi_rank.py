"""Two-rank GPU parity of the sharded path (skipped on a box with fewer than two GPUs): tools/mgpu_check.py under
torch.distributed.run -- every rank linearizes its edge shard, the partials are combined over NCCL in every reduce mode
(full allreduce / overlap-only allreduce (+ allgather) / ownership-based: landmark blocks summed, owned pose ranges gathered),
and H, b, chi2, the over-threshold counts and the state after a full step must equal the single-GPU result on every rank."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("size", [(3000, 700, 30000), (20000, 5000, 200000)])
def test_two_ranks_sharded_build_and_step_equal_single_gpu(built_lib, size):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mgpu_check.py")] + [str(x) for x in size]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "MGPU_CHECK OK world 2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_peer_modes_need_the_peer_mappings(built_lib):
    """reduce_mode 4 / 5 (combines over NVLink peer memory) are refused until the ranks' value buffers have been exchanged and opened;
    the export half works on a single GPU (handle + offset of the one allocation that holds values, statistics and barrier slots)."""
    from helpers import synth_problem
    from prb_project_bearing_only_slam_b200 import capi
    _, pr = synth_problem(300, 80, 3000, seed=5)
    ctx = capi.Context(solver=capi.SOLVER_PCG)
    with pytest.raises(capi.BosError):
        ctx.peer_export()                       # nothing uploaded yet
    pr.upload(ctx)
    for mode in (4, 5):
        with pytest.raises(capi.BosError):
            ctx.set_reduce_mode(mode)
    handle, offset = ctx.peer_export()
    assert len(handle) == capi.IPC_HANDLE_BYTES and offset >= 0 and any(handle)
    with pytest.raises(capi.BosError):
        ctx.peer_open([handle], [offset])       # a single rank has no peers
    ctx.set_reduce_mode(3)
    ctx.close()
