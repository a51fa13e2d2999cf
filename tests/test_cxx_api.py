"""The C++ host layer (include/gridpp.h): a user program written against the reference's C++ API (namespace gridpp,
std::vector types, value returns, std::invalid_argument / std::runtime_error) is compiled with g++ against the header,
linked with libgridpp_b200.so and run.

CPU: the program builds, links, and its host-side behaviour (containers, validation order, exception types, the
no-device failure) holds. GPU: everything it computes equals what the Python mirror gets through the same C ABI bit
for bit, and the oracle within the 1e-5 bar.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import bindings as B
from util import assert_bit_exact, assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "gridpp_b200")
f32 = np.float32


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cxx") / "cxx_api_driver")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "cxx", "cxx_api_driver.cpp"),
           "-L", LIBDIR, "-lgridpp_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cxx_program_builds_and_host_checks_pass(driver):
    res = subprocess.run([driver, "checks"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "checks ok" in res.stdout


@pytest.mark.gpu
def test_cxx_api_matches_mirror_and_oracle(driver, gpp, orc, tmp_path):
    rng = np.random.default_rng(77)
    ny, nx, S, E, hw, T, mp = 48, 64, 150, 6, 3, 9, 12
    dx = 1000.0
    y, x = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    y, x = y.astype(f32), x.astype(f32)
    gelev, glaf = rng.uniform(0, 400, (ny, nx)).astype(f32), rng.uniform(0, 1, (ny, nx)).astype(f32)
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    pelev, plaf = rng.uniform(0, 400, S).astype(f32), rng.uniform(0, 1, S).astype(f32)
    bg = (rng.normal(size=(ny, nx)) * 3).astype(f32)
    bg[5, 7] = np.nan
    bvar = rng.uniform(0.5, 2, (ny, nx)).astype(f32)
    ens = (bg[:, :, None] + rng.normal(size=(ny, nx, E))).astype(f32)
    ens[5, 7, :] = 0
    obs = rng.normal(size=S).astype(f32)
    ratios, sigmas = np.full(S, 0.5, f32), np.full(S, 0.7, f32)
    thr = np.linspace(-6, 6, T).astype(f32)
    qfield = rng.uniform(0, 1, (ny, nx)).astype(f32)
    qlat, qlon = rng.uniform(0, ny * dx, 20).astype(f32), rng.uniform(0, nx * dx, 20).astype(f32)
    h, v, w, cv_dist, radius, quantile = 8000.0, 300.0, 0.6, 1500.0, 9000.0, 0.8
    arrays = dict(glats=y, glons=x, gelevs=gelev, glafs=glaf, plats=py, plons=px, pelevs=pelev, plafs=plaf, background=bg, bvariance=bvar,
                  ensemble=ens, obs=obs, ratios=ratios, sigmas=sigmas, thresholds=thr, quantile_field=qfield, qlats=qlat, qlons=qlon)
    for name, a in arrays.items():
        np.ascontiguousarray(a, f32).tofile(str(tmp_path / ("in_%s.bin" % name)))
    meta = dict(ny=ny, nx=nx, nS=S, nE=E, halfwidth=hw, max_points=mp, ctype=1, quantile=quantile, h=h, v=v, w=w, cv_dist=cv_dist, radius=radius)
    (tmp_path / "meta.txt").write_text("".join("%s %r\n" % kv for kv in meta.items()))
    res = subprocess.run([driver, "run", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "run ok" in res.stdout, res.stdout + res.stderr

    def out(name, shape=None, dtype=f32):
        a = np.fromfile(str(tmp_path / ("out_%s.bin" % name)), dtype=dtype)
        return a.reshape(shape) if shape is not None else a

    # ---- the Python mirror through the same C ABI: identical bits
    grid = gpp.Grid(y, x, gelev, glaf, gpp.Cartesian)
    points = gpp.Points(py, px, pelev, plaf, gpp.Cartesian)
    s = gpp.BarnesStructure(h, v, w)
    pbg = gpp.nearest(grid, points, bg)
    assert_bit_exact(out("pbackground"), pbg, "nearest(grid, points, field)")
    oi = gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, s, mp)
    assert_bit_exact(out("oi", (ny, nx)), oi, "optimal_interpolation(Grid)")
    full, var = gpp.optimal_interpolation_full(grid, bg, bvar, points, obs, ratios, pbg, gpp.nearest(grid, points, bvar), s, mp)
    assert_bit_exact(out("full", (ny, nx)), full, "optimal_interpolation_full")
    assert_bit_exact(out("full_variance", (ny, nx)), var, "analysis variance")
    cv = gpp.CrossValidation(gpp.MultipleStructure(gpp.BarnesStructure(h), gpp.LinearStructure(0, 0.2, 0), gpp.CressmanStructure(0, 0, 0.7)), cv_dist)
    oi_cv = gpp.optimal_interpolation(grid.to_points(), bg.ravel(), points, obs, ratios, pbg, cv, mp, False)
    assert_bit_exact(out("oi_cv"), oi_cv, "optimal_interpolation(Points, CrossValidation(MultipleStructure))")
    pens = np.stack([gpp.nearest(grid, points, ens[:, :, e]) for e in range(E)], axis=1)
    assert_bit_exact(out("pensemble", (S, E)), pens, "nearest(grid, points, vec3)")
    ensi = gpp.optimal_interpolation_ensi(grid, ens, points, obs, sigmas, pens, s, mp)
    assert_bit_exact(out("ensi", (ny, nx, E)), ensi, "optimal_interpolation_ensi")
    assert_bit_exact(out("search", (ny, nx)), gpp.neighbourhood_search(bg, bvar, 2, 1.5, 2.0, 0.2), "neighbourhood_search")
    assert_bit_exact(out("gradient_minmax", (ny, nx)), gpp.calc_gradient(bvar, bg, gpp.MinMax, 2, 3, 0.1, -1.0), "calc_gradient MinMax")
    assert_bit_exact(out("gradient_regression", (ny, nx)), gpp.calc_gradient(bvar, bg, gpp.LinearRegression, 2), "calc_gradient LinearRegression")
    pobs2 = (obs[:, None] + f32(0.125) * np.arange(E, dtype=f32)[None, :]).astype(f32)
    assert_bit_exact(out("ebesc", (ny, nx, E)), gpp.optimal_interpolation_ensi_multi_ebesc(grid, bvar, ens, points, pobs2, ratios, pens, s, mp, False),
                     "optimal_interpolation_ensi_multi_ebesc")
    assert_bit_exact(out("ebe", (ny, nx, E)), gpp.optimal_interpolation_ensi_multi_ebe(grid, bvar, ens, ens, points, pobs2, ratios, pens, pens, s, mp),
                     "optimal_interpolation_ensi_multi_ebe")
    assert_bit_exact(out("utem", (ny, nx, E)), gpp.optimal_interpolation_ensi_multi_utem(grid, bvar, ens, ens, points, obs, ratios, pens, pens, s, mp, False),
                     "optimal_interpolation_ensi_multi_utem")
    assert_bit_exact(out("staticcorr", (20, S)), gpp.staticcorr_points(gpp.Points(qlat, qlon, None, None, gpp.Cartesian), points, s, 5), "staticcorr_points")
    for name, stat in (("mean", gpp.Mean), ("min", gpp.Min), ("max", gpp.Max)):
        assert_bit_exact(out(name, (ny, nx)), gpp.neighbourhood(bg, hw, stat), "neighbourhood " + name)
    assert_bit_exact(out("ens_mean", (ny, nx)), gpp.neighbourhood(ens, hw, gpp.Mean), "neighbourhood(vec3)")
    assert_bit_exact(out("qf", (ny, nx)), gpp.neighbourhood_quantile_fast(bg, quantile, hw, thr), "quantile_fast")
    assert_bit_exact(out("qf_field", (ny, nx)), gpp.neighbourhood_quantile_fast(bg, qfield, hw, thr), "quantile_fast(vec2 quantile)")
    assert_bit_exact(out("qf_ens", (ny, nx)), gpp.neighbourhood_quantile_fast(ens, quantile, hw, thr), "quantile_fast(vec3)")
    assert_bit_exact(out("thresholds"), gpp.get_neighbourhood_thresholds(bg, T), "get_neighbourhood_thresholds")
    rad3 = np.full(S, f32(radius) / f32(3), f32)
    assert_bit_exact(out("gridding", (ny, nx)), gpp.gridding(grid, points, obs, radius, 2, gpp.Mean), "gridding")
    assert_bit_exact(out("gridding_nearest", (ny, nx)), gpp.gridding_nearest(grid, points, obs, 0, gpp.Max), "gridding_nearest")
    assert_bit_exact(out("count", (ny, nx)), gpp.count(points, grid, radius), "count")
    assert_bit_exact(out("distance"), gpp.distance(grid, points, 3), "distance")
    assert_bit_exact(out("fill", (ny, nx)), gpp.fill(grid, bg, points, rad3, -1.0, False), "fill")
    assert_bit_exact(out("fill_missing", (ny, nx)), gpp.fill_missing(bg), "fill_missing")
    assert_bit_exact(out("doping_circle", (ny, nx)), gpp.doping_circle(grid, bg, points, obs, rad3, 100.0), "doping_circle")
    assert_bit_exact(out("doping_square", (ny, nx)), gpp.doping_square(grid, bg, points, obs, np.ones(S, np.int32)), "doping_square")
    assert_bit_exact(out("gridding", (ny, nx)), orc.gridding((y, x), (py, px), obs, radius, 2, B.MEAN, B.CARTESIAN), "C++ gridding vs oracle")
    assert_bit_exact(out("nbh_std", (ny, nx)), gpp.neighbourhood(bg, hw, gpp.Std), "neighbourhood Std")
    assert_bit_exact(out("nbh_median", (ny, nx)), gpp.neighbourhood(bg, hw, gpp.Median), "neighbourhood Median")
    assert_bit_exact(out("nbh_quantile", (ny, nx)), gpp.neighbourhood_quantile(bg, quantile, hw), "neighbourhood_quantile")
    assert_bit_exact(out("nbh_quantile_ens", (ny, nx)), gpp.neighbourhood_quantile(ens, quantile, hw), "neighbourhood_quantile(vec3)")
    assert_bit_exact(out("brute_variance", (ny, nx)), gpp.neighbourhood_brute_force(bg, hw, gpp.Variance), "neighbourhood_brute_force")
    assert_bit_exact(out("member_median"), gpp.calc_statistic(ens.reshape(-1, E), gpp.Median), "calc_statistic(vec2)")
    assert_bit_exact(out("member_q"), gpp.calc_quantile(ens.reshape(-1, E), quantile), "calc_quantile(vec2)")
    assert_bit_exact(out("member_q_field", (ny, nx)), gpp.calc_quantile(ens, qfield), "calc_quantile(vec3, vec2)")
    assert_bit_exact(out("interp"), gpp.interpolate(thr, thr, np.full(T, thr.sum(dtype=f32), f32) * 0 + gpp.calc_statistic(np.tile(thr, (T, 1)), gpp.Sum)), "interpolate")
    assert_bit_exact(out("nbh_median", (ny, nx)), orc.neighbourhood(bg, hw, B.MEDIAN), "C++ neighbourhood Median vs oracle")
    assert_bit_exact(out("nbh_quantile_ens", (ny, nx)), orc.neighbourhood_window(ens, hw, B.QUANTILE, quantile), "C++ neighbourhood_quantile(vec3) vs oracle")
    assert_bit_exact(out("nearest_grid", (ny, nx)), gpp.nearest(points, grid, obs), "nearest(points, grid, values)")
    nn = np.array([points.get_nearest_neighbour(a, b) for a, b in zip(qlat, qlon)], np.int32)
    assert_bit_exact(out("nn", dtype=np.int32), nn, "get_nearest_neighbour")
    found = [points.get_neighbours_with_distance(a, b, radius) for a, b in zip(qlat, qlon)]
    assert_bit_exact(out("counts", dtype=np.int32), np.array([len(i) for i, _ in found], np.int32), "get_num_neighbours")
    assert_bit_exact(out("neighbours", dtype=np.int32), np.concatenate([i for i, _ in found]).astype(np.int32), "get_neighbours")
    assert_bit_exact(out("distances"), np.concatenate([d for _, d in found]).astype(f32), "get_neighbours_with_distance")
    closest = np.concatenate([points.get_closest_neighbours(a, b, 5) for a, b in zip(qlat, qlon)]).astype(np.int32)
    assert_bit_exact(out("closest", dtype=np.int32), closest, "get_closest_neighbours")
    grid_nn = np.concatenate([grid.get_nearest_neighbour(a, b) for a, b in zip(qlat, qlon)]).astype(np.int32)
    assert_bit_exact(out("grid_nn", dtype=np.int32), grid_nn, "Grid::get_nearest_neighbour")
    assert_bit_exact(out("grid_neighbours", dtype=np.int32), grid.get_neighbours(qlat[0], qlon[0], radius).ravel().astype(np.int32), "Grid::get_neighbours")
    px5, py5, pz5 = points._set.xyz()
    p5 = np.stack([px5, py5, pz5, pelev, plaf], axis=1).astype(f32)
    first = np.repeat(p5[:1], S - 1, axis=0)
    assert_bit_exact(out("corr"), s.corr(first, p5[1:]), "BarnesStructure::corr")
    assert_bit_exact(out("corr_cv_background"), cv.corr_background(first, p5[1:]), "CrossValidation::corr_background")

    # ---- and the oracle, at the bar of the parity tests
    bpts, opts = (y, x, gelev, glaf), (py, px, pelev, plaf)
    want = orc.optimal_interpolation(bpts, bg, opts, obs, ratios, pbg, B.make_structure(B.BARNES, h, v, w), mp, B.CARTESIAN)
    assert_close(out("oi"), want, 3.0, 1e-5, "C++ optimal_interpolation vs oracle")
    assert_close(out("mean", (ny, nx)), orc.neighbourhood(bg, hw, B.MEAN), 3.0, 1e-5, "C++ neighbourhood mean vs oracle")
    assert_bit_exact(out("max", (ny, nx)), orc.neighbourhood(bg, hw, B.MAX), "C++ neighbourhood max vs oracle")
    assert_close(out("qf", (ny, nx)), orc.neighbourhood_quantile_fast(bg, quantile, hw, thr), 6.0, 1e-5, "C++ quantile_fast vs oracle")
    assert_bit_exact(nn, orc.points_nearest(py, px, B.CARTESIAN, qlat, qlon).astype(np.int32), "nearest index vs oracle")
