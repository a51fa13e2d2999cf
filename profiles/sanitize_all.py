"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_all.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp

rng = np.random.default_rng(0)
f32 = np.float32
# neighbourhood: TMA kernels (static and run-time half-widths), plain kernels, missing values
for shape in ((70, 256), (40, 300), (33, 77)):
    f = rng.uniform(size=shape).astype(f32)
    f[rng.uniform(size=shape) < 0.03] = np.nan
    for hw in (1, 7, 9, 15):
        for st in (gpp.Mean, gpp.Count, gpp.Min, gpp.Max):
            gpp.neighbourhood(f, hw, st)
    thr = np.linspace(0, 1, 20).astype(f32)
    for hw in (2, 15):
        gpp.neighbourhood_quantile_fast(f, 0.5, hw, thr)
        gpp.neighbourhood_quantile_fast(f, rng.uniform(size=shape).astype(f32), hw, thr[:7])
    gpp.neighbourhood_quantile_fast(f, 0.5, 3, [3, 1, 2])          # general kernel
gpp.neighbourhood(rng.uniform(size=(20, 256, 3)).astype(f32), 2, gpp.Mean)
gpp.neighbourhood_quantile_fast(rng.uniform(size=(20, 256, 3)).astype(f32), 0.5, 2, [0.2, 0.5, 0.8])
gpp.get_neighbourhood_thresholds(rng.uniform(size=(50, 60)).astype(f32), 7)
# OI: register kernel (grid tiles and points runs), Cholesky kernels, general kernel, spatial structure, EnSI
ny, nx, dx = 24, 40, 250.0
y, x = np.meshgrid(20000 + np.arange(ny) * dx, 30000 + np.arange(nx) * dx, indexing="ij")
S = 900
py, px = rng.uniform(-20000, 70000, S).astype(f32), rng.uniform(-10000, 90000, S).astype(f32)
bg = rng.normal(size=(ny, nx)).astype(f32)
pbg = rng.normal(size=S).astype(f32)
obs = (pbg + rng.normal(size=S)).astype(f32)
ratios = np.full(S, 0.5, f32)
grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
for mp in (30, 50, 100, 0):
    gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, gpp.BarnesStructure(10000), mp)
gpp.optimal_interpolation_full(grid, bg, np.ones(bg.shape), points, obs, ratios, pbg, np.ones(S), gpp.BarnesStructure(10000), 50)
gpp.optimal_interpolation(gpp.Points(y.ravel(), x.ravel(), type=gpp.Cartesian), bg.ravel(), points, obs, ratios, pbg, gpp.BarnesStructure(10000), 30)
be, pe = rng.uniform(0, 100, y.shape).astype(f32), rng.uniform(0, 100, S).astype(f32)
gpp.optimal_interpolation(gpp.Grid(y, x, be, type=gpp.Cartesian), bg, gpp.Points(py, px, pe, type=gpp.Cartesian), obs, ratios, pbg,
                          gpp.SoarStructure(10000, 200), 20)
sg = gpp.Grid(*np.meshgrid(np.linspace(0, 40000, 4), np.linspace(0, 50000, 5), indexing="ij"), type=gpp.Cartesian)
gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, gpp.BarnesStructure(sg, np.full((4, 5), 8000.0), np.zeros((4, 5)), np.zeros((4, 5))), 20)
E = 8
gpp.optimal_interpolation_ensi(grid, rng.normal(size=(ny, nx, E)).astype(f32), points, obs, ratios, rng.normal(size=(S, E)).astype(f32),
                               gpp.BarnesStructure(10000), 40)
idx = grid.get_nearest_neighbour(25000.0, 35000.0)
points.get_neighbours(25000.0, 35000.0, 5000.0)
gpp.nearest(grid, points, bg)
gpp.synchronize()
print("sanitize_all: done")
