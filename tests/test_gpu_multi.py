"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): scripts/mgpu_check.py under
torch.distributed.run — observation shards (NCCL all-reduce of g and H per Newton iteration) and the node group
(quadrature nodes, sample blocks and prediction rows split over the ranks) against the same calls on one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least two GPUs")
def test_sharded_paths_match_single_gpu():
    n = min(_ngpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "scripts", "mgpu_check.py"), "200000", "60"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stderr[-4000:]
    assert "MGPU_CHECK PASS" in out.stdout, out.stdout[-4000:]
