"""GPU, >= 2 devices (skipped on a single-GPU box): the real multi-rank paths — sharded store search
(NCCL all-gather / all-to-all exchange + merge kernel, and the peer-memory exchange kernel) and the
sharded distributed loss — launched as torchrun world-size-2 jobs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, *args):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines, dec = [], json.JSONDecoder()
    for l in p.stdout.splitlines():          # tolerate two ranks' objects landing on one line
        i = l.find("{")
        while i >= 0:
            try:
                obj, end = dec.raw_decode(l, i)
            except ValueError:
                break
            lines.append(obj)
            i = l.find("{", end)
    assert p.returncode == 0 and len(lines) == 2, p.stdout[-2000:] + p.stderr[-2000:]
    return lines


def test_sharded_search_two_ranks_nccl():
    for r in _torchrun("dist_search_check.py", 29811):
        assert r["ok"], r


def test_sharded_distributed_loss_two_ranks_nccl():
    for r in _torchrun("dist_loss_check.py", 29812):
        assert r["ok"], r


def test_peer_memory_exchange_stress_two_ranks():
    """Exchange + merge as one kernel over peer-mapped memory: 80 searches with changing (Q, k)
    (buffer reuse and regrowth), every result bit-identical to a single full index."""
    for r in _torchrun("peer_stress.py", 29813, "80"):
        assert r["ok"] and r["mismatches"] == 0 and r["used_peer"] == 80, r
