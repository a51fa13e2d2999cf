"""BASELINE.json config 5 (EnSI: 2500 x 2500 grid, dx 200 m, 20 members, 5000 observations, Barnes 10 km,
max_points 50) through the host API, with GPP_TRACE phase timings, next to the reference's (serial) CPU path on a
row-strided sample. usage: python profiles/ensi_time.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GPP_TRACE", "1")
import gridpp_b200 as gpp
from oracle import bindings as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
dx, E, S = 200.0, 20, 5000
rng = np.random.default_rng(1000)
ext = 2500 * dx
y, x = np.meshgrid(np.arange(n, dtype=np.float32) * (ext / n), np.arange(n, dtype=np.float32) * (ext / n), indexing="ij")
py, px = (rng.random(S) * ext).astype(np.float32), (rng.random(S) * ext).astype(np.float32)
bg = (rng.standard_normal((n, n, 1), dtype=np.float32) * 2 + rng.standard_normal((n, n, E), dtype=np.float32)).astype(np.float32)
pbg = rng.standard_normal((S, E)).astype(np.float32)
obs = rng.standard_normal(S).astype(np.float32)
sig = np.full(S, 0.5, np.float32)
grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
s = gpp.BarnesStructure(10000)
for it in range(3):
    t0 = time.perf_counter()
    out = gpp.optimal_interpolation_ensi(grid, bg, points, obs, sig, pbg, s, 50)
    dt = time.perf_counter() - t0
    print("EnSI host call %d: %.1f ms  (%.2f M gridpoints/s end to end)" % (it, 1e3 * dt, n * n / dt / 1e6), flush=True)
kind = "ref" if B.available("ref") else "oracle"
lib = B.load(kind)
pick = np.arange(0, n * n, max(1, n * n // 4000))[:4000]
timing = []
want = lib.optimal_interpolation_ensi((y.ravel()[pick], x.ravel()[pick], None, None), bg.reshape(-1, E)[pick], (py, px, None, None), obs, sig,
                                      pbg, B.make_structure(B.BARNES, 10000.0), 50, B.CARTESIAN, timing=timing)
err = np.abs(out.reshape(-1, E)[pick] - want) / np.maximum(np.abs(want), 2.0)
print("CPU %s (serial code, oi_ensi.cpp:203-206): %d gridpoints in %.2f s = %.0f gridpoints/s; parity max rel err %.2e" % (
    kind, pick.size, timing[0], pick.size / timing[0], err.max()))
