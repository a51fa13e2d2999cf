"""End-to-end times (host API, numpy in / numpy out) of the SURVEY 8(f) rows next to the compiled reference on the same inputs
and the box's host threads, with the two results compared. usage: python profiles/next_rows_time.py [n]   (n x n grid, default 2000)
Writes one JSON object to stdout."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from oracle import bindings as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
f32 = np.float32
rng = np.random.default_rng(5)
dx = 250.0
y, x = np.meshgrid(np.arange(n, dtype=f32) * dx, np.arange(n, dtype=f32) * dx, indexing="ij")
S = 10000
py, px = rng.uniform(0, n * dx, S).astype(f32), rng.uniform(0, n * dx, S).astype(f32)
values = rng.standard_normal(S).astype(f32)
field = (rng.standard_normal((n, n)) * 3).astype(f32)
field[rng.uniform(size=field.shape) < 0.01] = np.nan
elev = (300 + 200 * rng.standard_normal((n, n))).astype(f32)
laf = rng.uniform(0, 1, (n, n)).astype(f32)
grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
kind = "ref" if B.available("ref") else "oracle"
cpu = B.load(kind)
threads = os.cpu_count() or 1
cpu.set_omp_threads(threads)
G, P = (y, x), (py, px)
radius = 5000.0
radii = np.full(S, 2000.0, f32)


def timed(fn, reps=2):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t.append(time.perf_counter() - t0)
    return min(t), out


cases = [
    ("gridding mean, radius 5 km", lambda: gpp.gridding(grid, points, values, radius, 1, gpp.Mean), lambda: cpu.gridding(G, P, values, radius, 1, B.MEAN, B.CARTESIAN), 0),
    ("gridding median, radius 5 km", lambda: gpp.gridding(grid, points, values, radius, 1, gpp.Median), lambda: cpu.gridding(G, P, values, radius, 1, B.MEDIAN, B.CARTESIAN), 0),
    ("gridding_nearest max", lambda: gpp.gridding_nearest(grid, points, values, 0, gpp.Max), lambda: cpu.gridding(G, P, values, 0, 0, B.MAX, B.CARTESIAN, nearest=True), 0),
    ("count(points -> grid), radius 5 km", lambda: gpp.count(points, grid, radius), lambda: cpu.count(P, G, radius, B.CARTESIAN), 0),
    ("distance(points -> grid), 3 nearest", lambda: gpp.distance(points, grid, 3), lambda: cpu.distance(P, G, 3, B.CARTESIAN), 0),
    ("fill, 10 k circles of 2 km", lambda: gpp.fill(grid, field, points, radii, -1.0, False), lambda: cpu.fill(y, x, field, py, px, radii, -1.0, 0, B.CARTESIAN), 0),
    ("fill_missing", lambda: gpp.fill_missing(field), lambda: cpu.fill_missing(field), 0),
    ("doping_circle, 10 k circles of 2 km", lambda: gpp.doping_circle(grid, field, points, values, radii), lambda: cpu.doping(y, x, None, field, py, px, None, values, radii, float("nan"), B.CARTESIAN, False), 0),
    ("neighbourhood std hw 3", lambda: gpp.neighbourhood(field, 3, gpp.Std), lambda: cpu.neighbourhood(field, 3, B.STD), 1e-4),
    ("neighbourhood median hw 3", lambda: gpp.neighbourhood(field, 3, gpp.Median), lambda: cpu.neighbourhood(field, 3, B.MEDIAN), 0),
    ("neighbourhood_quantile 0.9 hw 3", lambda: gpp.neighbourhood_quantile(field, 0.9, 3), lambda: cpu.neighbourhood_window(field, 3, B.QUANTILE, 0.9), 0),
    ("neighbourhood_search hw 3", lambda: gpp.neighbourhood_search(field, laf, 3, 0.9, 1.0, 0.1), lambda: cpu.neighbourhood_search(field, laf, 3, 0.9, 1.0, 0.1), 0),
    ("calc_gradient MinMax hw 3", lambda: gpp.calc_gradient(elev, field, gpp.MinMax, 3), lambda: cpu.calc_gradient(elev, field, 0, 3), 0),
    ("calc_gradient LinearRegression hw 3", lambda: gpp.calc_gradient(elev, field, gpp.LinearRegression, 3), lambda: cpu.calc_gradient(elev, field, 10, 3), 1e-3),
]
out = {"grid": "%d x %d, dx %g m, %d observations" % (n, n, dx, S), "cpu": {"kind": "reference" if kind == "ref" else "port", "threads": threads}, "rows": {}}
for name, g_fn, c_fn, tol in cases:
    tg, got = timed(g_fn)
    t0 = time.perf_counter()
    want = c_fn()
    tc = time.perf_counter() - t0
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    same_nan = bool((np.isnan(got) == np.isnan(want)).mean() > 0.9999)
    ok = ~np.isnan(want) & ~np.isnan(got)
    err = float((np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1.0)).max()) if ok.any() else 0.0
    out["rows"][name] = {"gpu_s": round(tg, 5), "cpu_s": round(tc, 4), "speedup": round(tc / tg, 1), "max_rel_err": err, "equal": bool(err <= tol and same_nan)}
    print(name, out["rows"][name], file=sys.stderr, flush=True)
print(json.dumps(out))
