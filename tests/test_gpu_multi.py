"""Multi-GPU (NCCL) parity: the distributed transform equals the single-GPU transform.
Needs >= 2 GPUs; launched as subprocesses with torchrun-style env."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,no_p2p,gp", [(2, 0, "latbands"), (2, 1, "latbands"), (4, 0, "latbands"),
                                             (2, 0, "eq_regions"), (4, 0, "eq_regions"), (8, 0, "eq_regions")])
def test_distributed_equals_single(built, world, no_p2p, gp):
    """no_p2p = 0: GEMM / FFT epilogues write records into the consumer rank's buffer over NVLink (CUDA IPC),
    the transposition is only a barrier; no_p2p = 1: NCCL grouped send/recv all-to-all-v."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, ECT_NO_P2P=str(no_p2p), ECT_DIST_GP=gp)      # eq_regions: TRLTOG / TRGTOL as NCCL all-to-alls
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world + 10 * no_p2p + (20 if gp == "eq_regions" else 0)), os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK" in out.stdout


@pytest.mark.parametrize("world,V", [(2, 2), (4, 2), (8, 2)])
def test_vsets_equal_single(built, world, V):
    """NPRTRV > 1 (the reference benchmark's default 4 x 2 on 8 ranks): fields spread over V-sets, eq_regions grid-point
    tasks, TRLTOG / TRGTOL moving points and fields; results equal one rank's to rounding (field pairs of the FFT change)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, ECT_DIST_V=str(V))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29800 + world), os.path.join(ROOT, "tools", "dist_check_vsets.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "VSET_CHECK_OK" in out.stdout
