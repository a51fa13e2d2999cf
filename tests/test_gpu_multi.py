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


def test_single_process_multi_gpu_oi_equals_single_device(gpp):
    """gpp_optimal_interpolation_multi_gpu_host: one process, the rows of the grid split over the visible devices; bit-identical
    to the single-device call (analysis and, through optimal_interpolation_full's arguments, the variance path is the same code)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run on a box with >= 2 B200s: gpurun --gpus 2)")
    rng = np.random.default_rng(3)
    ny, nx, dx, S = 515, 640, 250.0, 2500
    y, x = np.meshgrid(np.arange(ny, dtype=np.float32) * dx, np.arange(nx, dtype=np.float32) * dx, indexing="ij")
    py, px = (rng.random(S) * ny * dx).astype(np.float32), (rng.random(S) * nx * dx).astype(np.float32)
    bg = rng.standard_normal((ny, nx)).astype(np.float32)
    pbg = rng.standard_normal(S).astype(np.float32)
    obs = (pbg + rng.standard_normal(S) * 0.5).astype(np.float32)
    ratios = np.full(S, 0.5, np.float32)
    grid, points, s = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    one = gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, s, 30)
    for nd in (0, 2):
        many = gpp.optimal_interpolation_multi_gpu(grid, bg, points, obs, ratios, pbg, s, 30, n_devices=nd)
        assert np.array_equal(one, many, equal_nan=True), "n_devices=%d" % nd
    pts = gpp.Points(y.ravel()[:100001], x.ravel()[:100001], type=gpp.Cartesian)      # a point set: split by points
    a = gpp.optimal_interpolation(pts, bg.ravel()[:100001], points, obs, ratios, pbg, s, 30)
    b = gpp.optimal_interpolation_multi_gpu(pts, bg.ravel()[:100001], points, obs, ratios, pbg, s, 30)
    assert np.array_equal(a, b, equal_nan=True)
