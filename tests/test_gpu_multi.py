"""Multi-GPU path on real GPUs (skipped with fewer than 2): env-index shards, no data-path collective, one NCCL
all-reduce of the stats block == the unsharded run (BASELINE config 4 with a SafetyWrapper band)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_rollout_nccl_allreduce_matches_unsharded():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-8000:]
    assert f"NCCL_SHARDED_OK {world}" in out.stdout, out.stdout[-2000:]
