"""ebesc / ebe / utem on config 5's geometry (2500 x 2500 x 20 members, 5000 observations, max_points 30), end to end through the
host API, with the reference's serial loops on a row-strided sample of the same points. usage: python profiles/ensi_multi_time.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from oracle import bindings as B

w = bench.ensi_inputs(0, bench.ENSI_N)
n, E, S, mp = bench.ENSI_N, bench.ENSI_E, bench.ENSI_S, 30
rng = np.random.default_rng(7)
grid, points = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian), gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s, so = gpp.BarnesStructure(bench.H_SCALE), B.make_structure(B.BARNES, bench.H_SCALE)
pobs2 = (w["obs"][:, None] + 0.1 * np.arange(E, dtype=np.float32)[None, :]).astype(np.float32)
pr, br = np.full(S, 0.25, np.float32), np.ones((n, n), np.float32)
bgc = w["bg"][::-1].copy()                      # another ensemble of the same shape for the correlations
pbgc = rng.standard_normal((S, E)).astype(np.float32)
cpu = B.load("ref" if B.available("ref") else "oracle")
calls = {
    "ebesc": lambda: gpp.optimal_interpolation_ensi_multi_ebesc(grid, br, w["bg"], points, pobs2, pr, w["pbg"], s, mp, False),
    "ebe": lambda: gpp.optimal_interpolation_ensi_multi_ebe(grid, br, w["bg"], bgc, points, pobs2, pr, w["pbg"], pbgc, s, mp, False),
    "utem": lambda: gpp.optimal_interpolation_ensi_multi_utem(grid, br, w["bg"], bgc, points, w["obs"], pr, w["pbg"], pbgc, s, mp, False),
}
out = {}
m = 40000
pick = np.arange(0, n * n, n * n // m)[:m]
bp = (w["y"].ravel()[pick], w["x"].ravel()[pick], None, None)
for kind, fn in calls.items():
    fn()
    t0 = time.perf_counter()
    got = fn()
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = cpu.ensi_multi(kind, bp, br.ravel()[pick], w["bg"].reshape(-1, E)[pick], bgc.reshape(-1, E)[pick] if kind != "ebesc" else None,
                          (w["py"], w["px"], None, None), w["obs"] if kind == "utem" else pobs2, pr, w["pbg"], pbgc if kind != "ebesc" else None, so, mp,
                          B.CARTESIAN, False)
    tc = time.perf_counter() - t0
    err = np.abs(got.reshape(-1, E)[pick] - want) / np.maximum(np.abs(want), 2.0)
    out[kind] = {"gpu_seconds_end_to_end": round(tg, 4), "gpu_gridpoints/s": round(n * n / tg), "cpu_gridpoints/s (serial, %d-point sample)" % m: round(m / tc),
                 "ratio": round(n * n / tg / (m / tc)), "max_rel_err": float(err.max()), "beyond_1e-5": int((err > 1e-5).sum())}
    print(kind, out[kind], file=sys.stderr, flush=True)
print(json.dumps(out))
