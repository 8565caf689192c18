"""Segment-split mode on real GPUs: one process per GPU under torchrun, checked against a
single handle holding the whole sequence (tests/mgpu_worker.py).  Needs >= 2 GPUs; the 1-GPU box skips it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,transport", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
def test_segment_split_matches_single_handle(world, transport):
    """transport: the per-sweep carries travel through peer mailboxes over NVLink (default) or NCCL all-gathers."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ, HML_EXCHANGE=transport, HML_EXPECT_TRANSPORT=transport)
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MGPU WORKER OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
