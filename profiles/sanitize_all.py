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
# ---- round 2 kernels
# OI with the analysis variance (cached inverse), a grid large enough for every unit size and for tile stealing
ny2, nx2 = 96, 128
y2, x2 = np.meshgrid(np.arange(ny2) * dx, np.arange(nx2) * dx, indexing="ij")
bg2 = rng.normal(size=(ny2, nx2)).astype(f32)
grid2 = gpp.Grid(y2, x2, type=gpp.Cartesian)
gpp.optimal_interpolation_full(grid2, bg2, np.ones(bg2.shape), points, obs, ratios, pbg, np.ones(S), gpp.BarnesStructure(10000), 30)
for _ in range(2):
    gpp.optimal_interpolation(grid2, bg2, points, obs, ratios, pbg, gpp.BarnesStructure(10000), 30)
# float mean kernel (hw 1, 2, 3, 5, 7 on rows that are a multiple of 4 and >= 256 wide), Std / Variance, the gather statistics
f = rng.uniform(size=(90, 512)).astype(f32)
for hw in (1, 2, 3, 5, 7):
    gpp.neighbourhood(f, hw, gpp.Mean)
    gpp.neighbourhood(f, hw, gpp.Sum)
f[rng.uniform(size=f.shape) < 0.02] = np.nan
for st in (gpp.Mean, gpp.Std, gpp.Variance, gpp.Median, gpp.RandomChoice):
    gpp.neighbourhood(f, 3, st)
gpp.neighbourhood_quantile(f, 0.3, 2)
gpp.neighbourhood_quantile(rng.uniform(size=(20, 30, 4)).astype(f32), 0.7, 1)
gpp.neighbourhood_brute_force(f[:30, :40], 2, gpp.Max)
rows = rng.normal(size=(500, 13)).astype(f32)
for st in (gpp.Mean, gpp.Min, gpp.Median, gpp.Max, gpp.Std, gpp.Variance, gpp.Sum, gpp.Count):
    gpp.calc_statistic(rows, st)
gpp.calc_quantile(rows, 0.9)
gpp.interpolate(rng.uniform(size=100).astype(f32), [0.0, 0.5, 1.0], [1.0, 2.0, 4.0])
# radius-query consumers
vals = rng.normal(size=S).astype(f32)
for st in (gpp.Mean, gpp.Median, gpp.Std, gpp.Count):
    gpp.gridding(grid, points, vals, 3000.0, 1, st)
gpp.gridding_nearest(grid, points, vals, 0, gpp.Max)
gpp.count(points, grid, 3000.0)
gpp.count(grid, points, 3000.0)
gpp.distance(grid, points, 3)
gpp.fill(grid, bg, points, np.full(S, 1500.0, f32), -1.0, False)
fm = bg.copy()
fm[rng.uniform(size=fm.shape) < 0.3] = np.nan
gpp.fill_missing(fm)
gpp.doping_circle(grid, bg, points, vals, np.full(S, 1500.0, f32))
gpp.doping_square(grid, bg, points, vals, np.full(S, 2, np.int32))
# ensi_multi (elimination kernel with and without the ensemble correlations), utem (EnSI kernel), staticcorr_points
nB = ny * nx
bgE, bgcE = rng.normal(size=(ny, nx, E)).astype(f32), rng.normal(size=(ny, nx, E)).astype(f32)
pbE, pbcE = rng.normal(size=(S, E)).astype(f32), rng.normal(size=(S, E)).astype(f32)
pobsE = (pbE + 0.5).astype(f32)
br = np.ones((ny, nx), f32)
for mp in (12, 100):
    gpp.optimal_interpolation_ensi_multi_ebesc(grid, br, bgE, points, pobsE, ratios, pbE, gpp.BarnesStructure(6000), mp, False)
    gpp.optimal_interpolation_ensi_multi_ebe(grid, br, bgE, bgcE, points, pobsE, ratios, pbE, pbcE, gpp.BarnesStructure(6000), mp)
gpp.optimal_interpolation_ensi_multi_utem(grid, br, bgE, bgcE, points, obs, ratios, pbE, pbcE, gpp.BarnesStructure(6000), 40, False)
gpp.staticcorr_points(gpp.Points(y.ravel()[:200], x.ravel()[:200], type=gpp.Cartesian), points, gpp.BarnesStructure(6000), 10)
gpp.staticcorr_points(gpp.Points(y.ravel()[:200], x.ravel()[:200], type=gpp.Cartesian), points, gpp.CressmanStructure(6000), 0)
gpp.synchronize()
print("sanitize_all: done")
