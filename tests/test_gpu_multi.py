"""
Two-GPU check of the sharded path (skipped on a single-GPU box): the finalize kernel's peer-memory
logging exchange against the NCCL all-reduce path and against the global mask counts.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["command_direction", "berkeley_humanoid"])
def test_peer_memory_logging_equals_nccl(config):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "8192", "30", config]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("PEER LOGGING OK") == 2, out.stdout[-3000:]
