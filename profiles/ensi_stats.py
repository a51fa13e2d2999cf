"""Jacobi statistics of the C5 EnSI analysis on an n x n sub-grid (needs a library built with -DENSI_STATS:
profiles/variants.sh estats "-DENSI_STATS"; GPP_B200_LIB=$PWD/scratch/lib_estats.so python profiles/ensi_stats.py [n])."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dx, E, S = 200.0, 20, 5000
rng = np.random.default_rng(1000)
ext = 2500 * dx
y, x = np.meshgrid(np.arange(n, dtype=np.float32) * dx, np.arange(n, dtype=np.float32) * dx, indexing="ij")
py, px = (rng.random(S) * ext).astype(np.float32), (rng.random(S) * ext).astype(np.float32)
bg = (rng.standard_normal((n, n, 1), dtype=np.float32) * 2 + rng.standard_normal((n, n, E), dtype=np.float32)).astype(np.float32)
pbg = rng.standard_normal((S, E)).astype(np.float32)
obs = rng.standard_normal(S).astype(np.float32)
sig = np.full(S, 0.5, np.float32)
grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
s = gpp.BarnesStructure(10000)
stats = (ctypes.c_ulonglong * 4)()
fn = _lib.lib.gpp_debug_ensi_stats
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
for it in range(2):
    fn(stats, 1)
    t0 = time.perf_counter()
    gpp.optimal_interpolation_ensi(grid, bg, points, obs, sig, pbg, s, 50)
    dt = time.perf_counter() - t0
    fn(stats, 0)
    pts = max(stats[0], 1)
    print("%d x %d: %.1f ms; points %d, warm starts %.1f%%, sweeps per point %.2f, rotations per point %.1f" % (
        n, n, 1e3 * dt, stats[0], 100.0 * stats[3] / pts, stats[1] / pts, stats[2] / pts), flush=True)
