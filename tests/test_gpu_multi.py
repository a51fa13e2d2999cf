"""Hardware parity of the N > 1 paths: spawns tests/mgpu_worker.py under torchrun with 2 ranks (NCCL) when the box has at
least two GPUs, and requires every sharded result to equal the single-GPU whole-field one."""
import json
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
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gpu_sharded_results_equal_whole_field():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run on a box with >= 2 B200s: gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert lines, "no result line; stderr tail: " + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["n_gpus"] == 2
    failed = [k for k, v in out["checks"].items() if not v]
    assert not failed and res.returncode == 0, "sharded != whole for: %s" % failed
    assert len(out["checks"]) >= 13
