"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): the row-sharded bank with the library's peer-memory
exchange and with an NCCL all-gather against a single-GPU scan of the whole bank — tools/sharded_check.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")
def test_sharded_bank_peer_and_nccl_exchange_equal_whole_bank_scan():
    n = min(_gpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["mismatches"] == 0 and res["peer_equals_nccl"] and res["world"] == n
