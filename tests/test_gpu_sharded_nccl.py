"""Real multi-GPU parity (world_size 2, scripts/check_sharded_nccl.py under torchrun): row-sharded K2 + ONE NCCL all-gather + K3 ==
the peer-memory push with a separate merge == the FUSED exchange (merge inside K2's last kernel) == pipelined submit / collect ==
unsharded search == CPU oracle; self-join pair counts add up across ranks.  Needs two GPUs on the box (skipped otherwise); the
single-GPU virtual-shard versions are tests/test_gpu_search.py::test_id_offset_and_virtual_shards_merge and
tests/test_gpu_multidevice.py (one-process sharded collection)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_search_over_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_sharded_nccl.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "SHARDED_NCCL_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
