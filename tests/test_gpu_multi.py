"""-m gpu (needs >= 2 GPUs, skipped otherwise): CFG/view-sharded step == single-GPU step (scripts/mgpu_check.py under
torchrun, NCCL)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4, 8])
def test_device_allgather_matches_nccl(world):
    """pf_allgather_views (push over NVLink into IPC-mapped buffers, flag handshake) == ncclAllGather, eager and as a
    replayed CUDA graph (scripts/allgather_check.py)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), str(ROOT / "scripts" / "allgather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout


@pytest.mark.parametrize("split_k", ["0", "1"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_single(world, split_k):
    """split_k 0: the sharded sampler result must be BIT-IDENTICAL to the single-GPU one; 1 (default kernels): within 2e-3."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import os
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), str(ROOT / "scripts" / "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, "PF_SPLIT_K": split_k})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


def test_transport_falls_back_to_nccl_together():
    """If the IPC / peer-access set-up of the device all-gather fails on ANY rank (forced here), every rank switches to NCCL at
    the same collective and the sharded result is unchanged."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    world = min(4, torch.cuda.device_count() // 2 * 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29650", str(ROOT / "scripts" / "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                       env={**os.environ, "PF_SPLIT_K": "0", "PF_FORCE_IPC_FAIL": "1", "MGPU_GRAPH_MODES": "1"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout and "fell back to NCCL" in r.stdout
