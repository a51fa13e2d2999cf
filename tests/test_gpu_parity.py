"""GPU parity tests: the CUDA path, called through the C ABI (via the gridpp_b200 mirror), against
  - the golden fixtures generated from the reference sources (tests/golden, see make_golden.py),
  - the plain-C oracle on fresh seeded inputs,
  - the reference's own known-answer values,
and, at the full sizes of BASELINE.json, through size-independent properties (subsample == oracle, shards ==
whole, locality of the stencil).

Bars (BASELINE.json north_star): index results bit-exact; fp32 fields within 1e-5 relative
(|got - want| <= 1e-5 * max(|want|, field scale)).
"""
import numpy as np
import pytest

from oracle import bindings as B
from util import assert_bit_exact, assert_close, assert_same_nan, golden, oracle_structure, parse_spec, product_structure

pytestmark = pytest.mark.gpu
f32 = np.float32
RTOL = 1e-5


# ------------------------------------------------------------------ structure functions ----------------
def test_structure_functions_golden(gpp):
    g = golden("structure")
    exact_families = ("barnes", "cressman", "powerlaw", "linear", "multiple")
    for name in g["names"]:
        name = str(name)
        s = product_structure(gpp, parse_spec(g[name + "__structure"]))
        for background, key in ((False, "__corr"), (True, "__corr_background")):
            got = s.corr_background(g["p1"], g["p2"]) if background else s.corr(g["p1"], g["p2"])
            if name.startswith(exact_families):
                assert_bit_exact(got, g[name + key], name + key)
            else:
                # Soar/Toar evaluate expf() in float in the reference; the device goes through the double exp
                assert_close(got, g[name + key], 1.0, 2e-7, name + key)
        assert f32(s.localization_distance()) == g[name + "__loc_dist"]


def test_barnes_known_answers(gpp):
    # tests/test_barnes_structure.py:8-33,85-97 of the reference: bit-exact float32 values
    x = [0, 1000, 2000, 3000, np.nan]
    p1 = np.zeros((5, 5), f32)
    p2 = np.zeros((5, 5), f32)
    p2[:, 1] = x
    barnes = gpp.BarnesStructure(2000)
    want = np.array([1, 0.8824968934059143, 0.6065306663513184, 0.32465246319770813, 0], f32)
    np.testing.assert_array_equal(barnes.corr(p1, p2), want)
    np.testing.assert_array_equal(barnes.corr(p2, p1), want)
    np.testing.assert_array_equal(gpp.CressmanStructure(2000).corr(p1, p2), np.array([1, 0.6, 0, 0, 0], f32))
    np.testing.assert_array_equal(gpp.CrossValidation(barnes, 1000).corr_background(p1, p2),
                                  np.array([0, 0, 0.6065306663513184, 0.32465246319770813, 0], f32))
    ans = {0: 1, 1000: 0.8824968934059143, 2000: 0.6065306663513184, 3000: 0.32465246319770813}
    for hmax in (0, 1000, 2000, 10000):
        s = gpp.BarnesStructure(2000, 0, 0, hmax)
        for dist, a in ans.items():
            got = s.corr([[0, 0, 0, 0, 0]], [[0, dist, 0, 0, 0]])[0]
            assert got == (f32(0) if dist > hmax else f32(a)), (hmax, dist, got)


# ------------------------------------------------------------------ index lookups (bit-exact) ----------
def test_index_queries_golden(gpp):
    g = golden("index_queries")
    for tname, t in (("geodetic", gpp.Geodetic), ("cartesian", gpp.Cartesian)):
        k = lambda n: g[tname + "__" + n]
        p = gpp.Points(k("lats"), k("lons"), type=t)
        assert_bit_exact(p._set.nearest(k("qlats"), k("qlons")), k("nearest"), tname + " nearest")
        assert_bit_exact(p._set.nearest(k("lats")[:200], k("lons")[:200], False), k("nearest_nomatch"), tname + " nearest nomatch")
        idx, dist, cnt = p._set.neighbours(k("qlats"), k("qlons"), float(k("radius")), capacity=96, with_distance=True)
        assert_bit_exact(cnt, k("nbr_count"), tname + " neighbour count")
        assert_bit_exact(idx, k("nbr_index"), tname + " neighbour index")
        assert_bit_exact(dist, k("nbr_dist"), tname + " neighbour distance")
        assert_bit_exact(p._set.closest(k("qlats"), k("qlons"), 5), k("closest5"), tname + " closest")
        out = gpp.nearest(p, gpp.Points(k("qlats"), k("qlons"), type=t), k("values"))
        assert_bit_exact(out, k("nearest_values"), tname + " nearest()")


def test_kdtree_known_answers(gpp):
    # tests/test_kdtree.py:8-53,149-161,188-203 of the reference
    tree = gpp.KDTree([60, 61, 62], [10, 10, 12])
    np.testing.assert_array_equal(tree.get_neighbours(60, 10, 1), [0])
    np.testing.assert_array_equal(tree.get_neighbours(60, 10, 112000), [0, 1])
    tree = gpp.KDTree([0, 1000, 2000], [0, 1000, 2000], gpp.Cartesian)
    _, dist = tree.get_neighbours_with_distance(100, 100, 1000)
    assert dist[0] == f32(100 * np.sqrt(2))
    tree = gpp.KDTree([50, 50, 51], [0, 0, 10])
    assert sorted(tree.get_neighbours(50, 0.001, 1000).tolist()) == [0, 1]
    assert sorted(tree.get_neighbours(50, 0, 1000).tolist()) == [0, 1]
    tree = gpp.KDTree([89, 89, 90, 90], [0, 180, 0, 10])
    idx, dist = tree.get_neighbours_with_distance(90, 0, 1000)
    assert sorted(idx.tolist()) == [2, 3] and np.abs(dist).max() < 1e-3
    points = gpp.Points([0, 1000, 2000], [0, 0, 0], [0, 0, 0], [0, 0, 0], gpp.Cartesian)
    np.testing.assert_array_equal(points.get_neighbours(900, 0, 501), [1])
    np.testing.assert_array_equal(points.get_neighbours(900, 0, 99.99), [])
    np.testing.assert_array_equal(points.get_neighbours(0, 0, 1000), [0])
    np.testing.assert_array_equal(points.get_neighbours(0, 0, 1001), [0, 1])
    np.testing.assert_array_equal(points.get_neighbours(0, 0, 1001, False), [1])
    assert points.get_num_neighbours(0, 0, 1001) == 2
    for lon in (-360, 0, 360):
        tree = gpp.KDTree([0], [lon])
        idx, dist = tree.get_neighbours_with_distance(0, 180, 1e9)
        assert idx[0] == 0 and abs(dist[0] - 12756274.0) < 2
    # empty sets: points.cpp:56-62, nearest.cpp:17-18
    assert gpp.Points().get_nearest_neighbour(0, 0) == -1
    assert gpp.Grid().get_nearest_neighbour(0, 0).size == 0
    assert np.isnan(gpp.nearest(gpp.Points(), gpp.Points([0, 1], [0, 1]), [])).all()
    lats, lons = np.meshgrid([0, 1, 2], [0, 1, 2], indexing="ij")
    grid = gpp.Grid(lats, lons)
    np.testing.assert_array_equal(grid.get_nearest_neighbour(10, 0.9), [2, 1])
    np.testing.assert_array_equal(gpp.nearest(grid, gpp.Points([10, -10], [0.9, 2.4]), np.arange(9).reshape(3, 3)), [7, 2])
    out = gpp.nearest(grid, grid, np.arange(18).reshape(2, 3, 3))
    np.testing.assert_array_equal(out, np.arange(18).reshape(2, 3, 3))


def test_nearest_large_random_vs_oracle(gpp, orc):
    rng = np.random.default_rng(11)
    for t, (lo, hi) in ((gpp.Cartesian, (0, 2e5)), (gpp.Geodetic, (40, 70))):
        la, lon = rng.uniform(lo, hi, 30000).astype(f32), rng.uniform(lo if t else -30, hi if t else 40, 30000).astype(f32)
        ql, qo = rng.uniform(lo, hi, 4000).astype(f32), rng.uniform(lo if t else -30, hi if t else 40, 4000).astype(f32)
        p = gpp.Points(la, lon, type=t)
        assert_bit_exact(p._set.nearest(ql, qo), orc.points_nearest(la, lon, t, ql, qo), "nearest type %d" % t)
    # a regular grid queried at its own nodes and at cell centres (exact ties -> lowest index)
    y, x = np.meshgrid(np.arange(60) * 1000.0, np.arange(50) * 1000.0, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    qy, qx = (y[:-1, :-1] + 500).ravel(), (x[:-1, :-1] + 500).ravel()
    assert_bit_exact(grid._set.nearest(qy, qx), orc.points_nearest(y, x, B.CARTESIAN, qy, qx), "tie handling")


# ------------------------------------------------------------------ optimal interpolation --------------
def test_oi_known_answers(gpp):
    # tests/test_optimal_interpolation.py:50-105,154-202 of the reference
    grid = gpp.Grid([[0, 0, 0]], [[0, 2500, 10000]], [[0, 0, 0]], [[0, 0, 0]], gpp.Cartesian)
    points = gpp.Points([0], [2500], [0], [0], gpp.Cartesian)
    s = gpp.BarnesStructure(2500)
    out = gpp.optimal_interpolation(grid, np.zeros([1, 3]), points, [1], [0.1], [0], s, 10)
    assert out.dtype == np.float32 and out.shape == (1, 3)
    np.testing.assert_array_almost_equal(out, [[np.exp(-0.5) / 1.1, 1 / 1.1, np.exp(-0.5 * 9) / 1.1]])
    out, var = gpp.optimal_interpolation_full(grid, np.zeros([1, 3]), np.ones([1, 3]), points, [1], [0.1], [0], [1], s, 10)
    assert abs(var[0, 1] - 0.1 / 1.1) < 1e-6
    bpoints = gpp.Points([0, 0, 0], [0, 2500, 10000], [0, 0, 0], [0, 0, 0], gpp.Cartesian)
    out, var = gpp.optimal_interpolation_full(bpoints, np.zeros(3), np.ones(3), points, [1], [0.1], [0], [1], s, 10)
    assert abs(var[1] - 0.1 / 1.1) < 1e-6
    # no observations -> background; NaN observations are ignored
    bg = np.arange(3, dtype=f32).reshape(1, 3)
    assert_bit_exact(gpp.optimal_interpolation(grid, bg, gpp.Points([], [], type=gpp.Cartesian), [], [], [], s, 10), bg)
    two = gpp.Points([0, 0], [2500, 5000], [0, 0], [0, 0], gpp.Cartesian)
    a = gpp.optimal_interpolation(grid, np.zeros([1, 3]), two, [1, np.nan], [0.1, 0.1], [0, 0], s, 10)
    b = gpp.optimal_interpolation(grid, np.zeros([1, 3]), points, [1], [0.1], [0], s, 10)
    assert_bit_exact(a, b)
    # extrapolation clamp: the increment never exceeds the largest innovation (:170-190)
    many = gpp.Points(np.zeros(5), np.arange(5) * 100.0 + 2000, type=gpp.Cartesian)
    out = gpp.optimal_interpolation(grid, np.zeros([1, 3]), many, [1, 0.9, 1, 0.95, 1], [0.01] * 5, [0] * 5, s, 10, False)
    assert out.max() <= 1.0


def _run_oi_case(gpp, g, name, bshape):
    spec = parse_spec(g[name + "__structure"])
    mp, extr, use_elev = (int(v) for v in g[name + "__args"])
    if use_elev:
        grid = gpp.Grid(g["y"], g["x"], g["belev"], g["blaf"], gpp.Cartesian)
        points = gpp.Points(g["py"], g["px"], g["pelev"], g["plaf"], gpp.Cartesian)
    else:
        grid = gpp.Grid(g["y"], g["x"], type=gpp.Cartesian)
        points = gpp.Points(g["py"], g["px"], type=gpp.Cartesian)
    s = product_structure(gpp, spec)
    return gpp.optimal_interpolation_full(grid, g["background"], np.ones(bshape, f32), points, g["pobs"], g["pratios"],
                                          g["pbackground"], np.ones(g["pobs"].size, f32), s, mp, bool(extr))


def test_oi_golden(gpp):
    g = golden("oi_c1_geodetic")
    grid = gpp.Grid(g["lats"], g["lons"])
    points = gpp.Points(g["plats"], g["plons"])
    # pbackground through the device nearest() must reproduce the fixture's indices (README.md:50 flow)
    assert_bit_exact(grid._set.nearest(g["plats"], g["plons"]), g["nearest_index"])
    assert_bit_exact(gpp.nearest(grid, points, g["background"]), g["pbackground"])
    s = product_structure(gpp, parse_spec(g["structure"]))
    out, var = gpp.optimal_interpolation_full(grid, g["background"], np.ones(g["background"].shape), points, g["pobs"], g["pratios"],
                                              g["pbackground"], np.ones(10), s, int(g["max_points"]))
    scale = float(np.nanstd(g["background"]))
    assert_close(out, g["analysis"], scale, RTOL, "C1 analysis")
    assert_close(var, g["analysis_variance"], 1.0, RTOL, "C1 variance")
    g = golden("oi_c3_density")
    scale = float(np.nanstd(g["background"]))
    for name in g["names"]:
        name = str(name)
        out, var = _run_oi_case(gpp, g, name, g["background"].shape)
        worst = assert_close(out, g[name + "__analysis"], scale, RTOL, name)
        assert_close(var, g[name + "__variance"], 1.0, RTOL, name + " variance")
        assert worst < RTOL


def test_oi_random_vs_oracle(gpp, orc):
    """BASELINE.json config 3 density (dx 250 m, Barnes h 10 km, ~42 candidates, max_points 30) on a sub-grid."""
    rng = np.random.default_rng(1000)
    ny, nx, dx = 120, 160, 250.0
    y, x = np.meshgrid(20000 + np.arange(ny) * dx, 30000 + np.arange(nx) * dx, indexing="ij")
    S = 1200   # 0.01 obs / km^2 over the 110 x 120 km area that can reach the sub-grid
    py, px = rng.uniform(-20000, 90000, S).astype(f32), rng.uniform(-10000, 110000, S).astype(f32)
    bg = (rng.normal(size=(ny, nx)) * 3).astype(f32)
    bg[rng.uniform(size=bg.shape) < 0.002] = np.nan
    pbg = rng.normal(size=S).astype(f32) * 3
    obs = (pbg + rng.normal(size=S) * 0.5).astype(f32)
    obs[rng.uniform(size=S) < 0.01] = np.nan
    ratios = np.full(S, 0.5, f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    for mp, extr in ((30, True), (30, False), (10, True), (1, True)):
        got, gvar = gpp.optimal_interpolation_full(grid, bg, np.ones(bg.shape), points, obs, ratios, pbg, np.ones(S), gpp.BarnesStructure(10000), mp, extr)
        want, wvar = orc.optimal_interpolation((y, x, None, None), bg, (py, px, None, None), obs, ratios, pbg,
                                               B.make_structure(B.BARNES, 10000.0), mp, B.CARTESIAN, allow_extrapolation=extr, want_variance=True)
        assert_close(got.ravel(), want, 3.0, RTOL, "OI mp=%d extr=%s" % (mp, extr))
        assert_close(gvar.ravel(), wvar, 1.0, RTOL, "OI variance mp=%d" % mp)
    # general path: more than 30 observations per point, unlimited, and a non-symmetric structure function
    sub = (slice(0, 40), slice(0, 50))
    gsub = gpp.Grid(y[sub], x[sub], type=gpp.Cartesian)
    for mp in (40, 0):
        got = gpp.optimal_interpolation(gsub, bg[sub], points, obs, ratios, pbg, gpp.BarnesStructure(10000), mp)
        want = orc.optimal_interpolation((y[sub], x[sub], None, None), bg[sub], (py, px, None, None), obs, ratios, pbg,
                                         B.make_structure(B.BARNES, 10000.0), mp, B.CARTESIAN)
        assert_close(got.ravel(), want, 3.0, RTOL, "OI general mp=%d" % mp)
    # 30 < max_points <= 64 with a symmetric structure function: the shared-memory Cholesky kernel (analysis, variance, clamp)
    for mp, extr, s_gpu, s_orc in ((31, True, gpp.BarnesStructure(10000), B.make_structure(B.BARNES, 10000.0)),
                                   (50, False, gpp.BarnesStructure(10000), B.make_structure(B.BARNES, 10000.0)),
                                   (64, True, gpp.PowerlawStructure(8000), B.make_structure(B.POWERLAW, 8000.0))):
        got, gvar = gpp.optimal_interpolation_full(gsub, bg[sub], np.ones(bg[sub].shape), points, obs, ratios, pbg, np.ones(S), s_gpu, mp, extr)
        want, wvar = orc.optimal_interpolation((y[sub], x[sub], None, None), bg[sub], (py, px, None, None), obs, ratios, pbg, s_orc, mp,
                                               B.CARTESIAN, allow_extrapolation=extr, want_variance=True)
        assert_close(got.ravel(), want, 3.0, RTOL, "OI Cholesky path mp=%d" % mp)
        assert_close(gvar.ravel(), wvar, 1.0, RTOL, "OI Cholesky path variance mp=%d" % mp)
    belev, pelev = rng.uniform(0, 300, y[sub].shape).astype(f32), rng.uniform(0, 300, S).astype(f32)
    got = gpp.optimal_interpolation(gpp.Grid(y[sub], x[sub], belev, type=gpp.Cartesian), bg[sub], gpp.Points(py, px, pelev, type=gpp.Cartesian),
                                    obs, ratios, pbg, gpp.CressmanStructure(30000, 200), 12)
    want = orc.optimal_interpolation((y[sub], x[sub], belev, None), bg[sub], (py, px, pelev, None), obs, ratios, pbg,
                                     B.make_structure(B.CRESSMAN, 30000.0, 200.0), 12, B.CARTESIAN)
    assert_close(got.ravel(), want, 3.0, RTOL, "OI non-symmetric Cressman")


def test_oi_device_api_and_row_shards(gpp, orc):
    """The device-resident entry point (what bench.py times and a multi-GPU driver shards by rows): any split of
    the background points into ranges gives exactly the whole-field result."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(5)
    ny, nx, dx = 64, 96, 1000.0
    y, x = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    S = 300
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    bg = rng.normal(size=(ny, nx)).astype(f32)
    pbg = rng.normal(size=S).astype(f32)
    obs = (pbg + rng.normal(size=S)).astype(f32)
    grid, points, s = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    state = gd.ObservationState(points, obs, np.full(S, 0.5, f32), pbg, s)
    d_bg = torch.from_numpy(bg.ravel()).cuda()
    whole = gd.optimal_interpolation(grid, d_bg, state, 30)
    parts = torch.full_like(d_bg, float("nan"))
    n = ny * nx
    cuts = [0, 1000, 1001, 3333, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        gd.optimal_interpolation(grid, d_bg, state, 30, out=parts, first=a, count=b - a)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    host = gpp.optimal_interpolation(grid, bg, points, obs, np.full(S, 0.5, f32), pbg, s, 30)
    assert_bit_exact(whole.cpu().numpy().reshape(ny, nx), host, "device vs host entry point")
    want = orc.optimal_interpolation((y, x, None, None), bg, (py, px, None, None), obs, np.full(S, 0.5, f32), pbg,
                                     B.make_structure(B.BARNES, 10000.0), 30, B.CARTESIAN)
    assert_close(host.ravel(), want, 1.0, RTOL, "device OI")


def test_oi_spatially_varying_structures_golden(gpp):
    """<Family>Structure(Grid, h, v, w, min_rho) (structure.cpp:168-184, :342, :492, :643, :790) through
    optimal_interpolation_full: analysis and analysis variance against fixtures generated from the compiled reference
    (tests/golden/make_golden_spatial.py). Also the reference's own spatial test, tests/test_barnes_structure.py:46-66."""
    g = golden("oi_spatial_structure")
    classes = {"barnes": gpp.BarnesStructure, "soar": gpp.SoarStructure, "toar": gpp.ToarStructure, "powerlaw": gpp.PowerlawStructure,
               "linear": gpp.LinearStructure}
    sgrid = gpp.Grid(g["gy"], g["gx"], type=gpp.Cartesian)
    bvar = np.ones(g["background"].shape, f32)
    for case in g["cases"]:
        name, _, elev, mp, extr, min_rho = str(case).split(",")
        elev, mp, extr, min_rho = int(elev), int(mp), int(extr), float(min_rho)
        grid = gpp.Grid(g["y"], g["x"], g["belev"], g["blaf"], gpp.Cartesian) if elev else gpp.Grid(g["y"], g["x"], type=gpp.Cartesian)
        points = gpp.Points(g["py"], g["px"], g["pelev"], g["plaf"], gpp.Cartesian) if elev else gpp.Points(g["py"], g["px"], type=gpp.Cartesian)
        s = classes[name](sgrid, g["h"], g["v"], g["w"], min_rho)
        key = "%s__elev%d__mp%d" % (name, elev, mp)
        out, var = gpp.optimal_interpolation_full(grid, g["background"], bvar, points, g["pobs"], g["pratios"], g["pbackground"],
                                                  np.ones(g["py"].size, f32), s, mp, bool(extr))
        # The observation-observation matrix of a spatially varying structure function is not symmetric
        # (corr(o_i, o_j) uses the scales at o_i); with Soar / Toar and elevations (signed differences,
        # structure.cpp:52-53,62-63) it is nearly singular at some grid points: analysis values up to 10^2 and
        # "variances" down to -160 against a background of order 1. The error there scales with the magnitude of the
        # field, so the parity metric uses the field's largest magnitude as its scale; at most 3 of the 1440 points may
        # exceed 1e-5 (none 1e-3).
        wa, wv = g[key + "__analysis"], g[key + "__variance"]
        sa, sv = max(2.0, float(np.abs(wa).max())), max(1.0, float(np.abs(wv).max()))
        assert_close(out, wa, sa, RTOL, "spatial " + key, allow_outliers=3)
        assert_close(out, wa, sa, 1e-3, "spatial (loose) " + key)
        assert_close(var, wv, sv, RTOL, "spatial variance " + key, allow_outliers=3)
        assert_close(var, wv, sv, 1e-3, "spatial variance (loose) " + key)
        out2 = gpp.optimal_interpolation(grid, g["background"], points, g["pobs"], g["pratios"], g["pbackground"], s, mp, bool(extr))
        assert_bit_exact(out2, out, "optimal_interpolation == _full for " + key)
    # tests/test_barnes_structure.py:46-66
    grid = gpp.Grid([[0, 0]], [[0, 2500]], [[0, 0]], [[0, 0]], gpp.Cartesian)
    s = gpp.BarnesStructure(grid, [[2500, 1]], [[0, 0]], [[0, 0]], 0.1)
    assert abs(s.localization_distance((0, 0)) - np.sqrt(-2 * np.log(0.1)) * 2500) < 1e-2
    assert abs(s.localization_distance((0, 2500)) - np.sqrt(-2 * np.log(0.1)) * 1) < 1e-4
    y, x = np.meshgrid(np.linspace(0, 1, 2), np.linspace(0, 1, 3))
    grid = gpp.Grid(y, x, y, y, gpp.Cartesian)
    valid = np.ones([3, 2])
    gpp.BarnesStructure(grid, valid, valid, valid)
    for inval in (np.ones([3, 4]), np.ones([2, 2]), np.ones([2, 4])):
        for args in ((inval, valid, valid), (valid, inval, valid), (valid, valid, inval)):
            with pytest.raises(ValueError):
                gpp.BarnesStructure(grid, *args)


# ------------------------------------------------------------------ neighbourhood filters --------------
STATS = {"mean": 0, "sum": 70, "count": 80, "min": 10, "max": 30}


def test_neighbourhood_golden(gpp):
    g = golden("neighbourhood")
    for hw in (0, 1, 7, 15, 70):
        for name, st in STATS.items():
            got = gpp.neighbourhood(g["field"], hw, st)
            want = g["hw%d__%s" % (hw, name)]
            assert got.dtype == np.float32
            if name in ("count", "min", "max"):
                assert_bit_exact(got, want, "hw=%d %s" % (hw, name))
            else:
                assert_close(got, want, 10.0 * (1 if name == "mean" else (2 * hw + 1) ** 2), 1e-6, "hw=%d %s" % (hw, name))


def test_neighbourhood_known_answers(gpp):
    # tests/test_neighbourhood.py:48-120,146-152 of the reference
    values = np.reshape(np.arange(25), [5, 5]).astype(f32)
    values[1, 3] = np.nan
    values[2, 4] = np.nan
    m = gpp.neighbourhood(values, 1, gpp.Mean)
    assert m[2, 2] == 12.5 and abs(m[0, 4] - 5.3333) < 1e-4
    assert (np.abs(gpp.neighbourhood(values, 100, gpp.Mean) - 12.086956) < 1e-4).all()
    assert_bit_exact(gpp.neighbourhood(values, 0, gpp.Mean), values)
    c = gpp.neighbourhood(values, 1, gpp.Count)
    assert c[2, 2] == 8 and c[0, 4] == 3 and (gpp.neighbourhood(values, 100, gpp.Count) == 23).all()
    mn, mx = gpp.neighbourhood(values, 1, gpp.Min), gpp.neighbourhood(values, 1, gpp.Max)
    assert mn[2, 2] == 6 and mn[0, 4] == 3 and mx[2, 2] == 18 and mx[0, 4] == 9
    assert (gpp.neighbourhood(values, 100, gpp.Min) == 0).all() and (gpp.neighbourhood(values, 100, gpp.Max) == 24).all()
    empty = np.zeros([5, 5], f32)
    empty[0:3, 0:3] = np.nan
    for st in (gpp.Mean, gpp.Min, gpp.Max, gpp.Sum):
        assert np.isnan(gpp.neighbourhood(empty, 1, st)[0:2, 0:2]).all()
    np.testing.assert_array_equal(gpp.neighbourhood(empty, 1, gpp.Count),
                                  [[0, 0, 2, 4, 4], [0, 0, 3, 6, 6], [2, 3, 5, 7, 6], [4, 6, 7, 8, 6], [4, 6, 6, 6, 4]])
    big = (np.arange(1, 1000, dtype=np.float64) ** 3).astype(f32)[:, None]
    np.testing.assert_array_almost_equal(gpp.neighbourhood(big, 0, gpp.Mean) / big - 1, np.zeros(big.shape), 6)
    # the median of the clipped window (neighbourhood.cpp:237-238 -> brute force -> calc_quantile(., 0.5))
    assert gpp.neighbourhood(values, 1, gpp.Median)[2, 2] == 12.5   # 8 valid values: half-way between 12 and 13


def test_neighbourhood_random_vs_oracle(gpp, orc):
    rng = np.random.default_rng(1000)
    for shape in ((257, 300), (64, 1000), (700, 31), (1, 1), (3, 500)):
        f = (rng.uniform(size=shape) * 10 - 3).astype(f32)
        f[rng.uniform(size=shape) < 0.02] = np.nan
        for hw in (0, 1, 7, 15, 64, 65):
            for name, st in STATS.items():
                got, want = gpp.neighbourhood(f, hw, st), orc.neighbourhood(f, hw, st)
                if name in ("count", "min", "max"):
                    assert_bit_exact(got, want, "%s hw=%d %s" % (shape, hw, name))
                else:
                    w = min(2 * hw + 1, shape[0]) * min(2 * hw + 1, shape[1])
                    assert_close(got, want, 10.0 * (1 if name == "mean" else w), 1e-6, "%s hw=%d %s" % (shape, hw, name))
    # high dynamic range: a running sum in float would lose the small values next to 1e9
    f = rng.uniform(size=(200, 200)).astype(f32)
    f[100, 100] = 1e9
    got, want = gpp.neighbourhood(f, 7, gpp.Mean), orc.neighbourhood(f, 7, gpp.Mean)
    far = np.ones(f.shape, bool)
    far[92:109, 92:109] = False
    assert_close(got[far], want[far], 1.0, 1e-6, "dynamic range (outside the spike's window)")
    assert_close(got, want, 1e9 / 225, 1e-6, "dynamic range (everywhere)")


def test_neighbourhood_row_tiles_equal_whole(gpp):
    """Row-tiled evaluation with halo rows (the multi-GPU decomposition) reproduces the whole-field result."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(9)
    ny, nx, hw = 300, 257, 7
    f = rng.uniform(size=(ny, nx)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.01] = np.nan
    d = torch.from_numpy(f).cuda()
    for st in (gpp.Mean, gpp.Min, gpp.Max, gpp.Count):
        whole = gd.neighbourhood(d, hw, st)
        tiles = []
        for r0, r1 in ((0, 100), (100, 101), (101, 300)):
            lo, hi = max(0, r0 - hw), min(ny, r1 + hw)
            tiles.append(gd.neighbourhood(d[lo:hi].contiguous(), hw, st, row0=r0 - lo, n_rows_out=r1 - r0))
        torch.cuda.synchronize()
        got = torch.cat(tiles)
        assert torch.equal(torch.nan_to_num(got, nan=-777.0), torch.nan_to_num(whole, nan=-777.0))
        assert_bit_exact(whole.cpu().numpy(), gpp.neighbourhood(f, hw, st))


def test_neighbourhood_copy_engine_path(gpp, orc):
    """Shapes the TMA-staged kernels take (row length a multiple of 4, >= 256): missing values, infinities (invalid,
    util.cpp:16-18), windows clipped on every edge, chunk/strip seams, and row tiles with halo (multi-GPU form)."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(77)
    for shape in ((300, 512), (37, 256), (1000, 260), (9, 1024)):
        f = (rng.uniform(size=shape) * 10 - 3).astype(f32)
        f[rng.uniform(size=shape) < 0.02] = np.nan
        f[rng.uniform(size=shape) < 0.002] = np.inf
        f[rng.uniform(size=shape) < 0.002] = -np.inf
        f[5:30, 100:140] = np.nan
        for hw in (1, 3, 7, 8, 12, 15, 31):
            for name, st in STATS.items():
                got, want = gpp.neighbourhood(f, hw, st), orc.neighbourhood(f, hw, st)
                if name in ("count", "min", "max"):
                    assert_bit_exact(got, want, "%s hw=%d %s" % (shape, hw, name))
                else:
                    w = min(2 * hw + 1, shape[0]) * min(2 * hw + 1, shape[1])
                    assert_close(got, want, 10.0 * (1 if name == "mean" else w), 1e-6, "%s hw=%d %s" % (shape, hw, name))
    # no missing values at all: the kernels' analytic-count path
    f = (rng.uniform(size=(500, 768)) * 10).astype(f32)
    for hw in (2, 7, 15):
        for name, st in STATS.items():
            got, want = gpp.neighbourhood(f, hw, st), orc.neighbourhood(f, hw, st)
            if name in ("count", "min", "max"):
                assert_bit_exact(got, want, "clean hw=%d %s" % (hw, name))
            else:
                assert_close(got, want, 10.0 * (1 if name == "mean" else (2 * hw + 1) ** 2), 1e-6, "clean hw=%d %s" % (hw, name))
    # high dynamic range next to ordinary values
    f = rng.uniform(size=(300, 400)).astype(f32)
    f[100, 100] = 1e9
    got, want = gpp.neighbourhood(f, 7, gpp.Mean), orc.neighbourhood(f, 7, gpp.Mean)
    far = np.ones(f.shape, bool)
    far[92:109, 92:109] = False
    assert_close(got[far], want[far], 1.0, 1e-6, "dynamic range (outside the spike's window)")
    assert_close(got, want, 1e9 / 225, 1e-6, "dynamic range (everywhere)")
    # row tiles with halo equal the whole field, bit for bit
    ny, nx, hw = 300, 260, 7
    f = rng.uniform(size=(ny, nx)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.01] = np.nan
    d = torch.from_numpy(f).cuda()
    for st in (gpp.Mean, gpp.Min, gpp.Max, gpp.Count):
        whole = gd.neighbourhood(d, hw, st)
        tiles = []
        for r0, r1 in ((0, 100), (100, 101), (101, 300)):
            lo, hi = max(0, r0 - hw), min(ny, r1 + hw)
            tiles.append(gd.neighbourhood(d[lo:hi].contiguous(), hw, st, row0=r0 - lo, n_rows_out=r1 - r0))
        torch.cuda.synchronize()
        got = torch.cat(tiles)
        assert torch.equal(torch.nan_to_num(got, nan=-777.0), torch.nan_to_num(whole, nan=-777.0))


def test_quantile_fast_golden(gpp):
    g = golden("quantile_fast")
    for hw in (1, 7, 15):
        for q in (0.0, 0.001, 0.5, 0.9, 0.999, 1.0):
            assert_bit_exact(gpp.neighbourhood_quantile_fast(g["field"], q, hw, g["thresholds"]), g["hw%d__q%g" % (hw, q)], "hw=%d q=%g" % (hw, q))
    assert_bit_exact(gpp.neighbourhood_quantile_fast(g["field"], g["qfield"], 3, g["thresholds"]), g["hw3__qfield"], "quantile field")
    assert_bit_exact(gpp.neighbourhood_quantile_fast(g["field"], 0.9, 2, g["thresholds_auto"]), g["hw2__auto_q0.9"], "auto thresholds")


def test_quantile_fast_known_answers(gpp, orc):
    # tests/test_neighbourhood_quantile_fast.py:34-85,129-136 of the reference
    values = np.reshape(np.arange(25), [5, 5]).astype(f32)
    values[1, 3] = np.nan
    values[2, 4] = np.nan
    field = np.reshape(np.arange(9), [3, 3]).astype(f32)
    for hw in (0, 1, 2):
        np.testing.assert_array_equal(gpp.neighbourhood_quantile_fast(field, 0.9, hw, [0]), np.zeros([3, 3]))
    thr = orc.get_neighbourhood_thresholds(values, 100)
    out = gpp.neighbourhood_quantile_fast(values, 0.5, 1, thr)
    assert out[2, 2] == 12 and out[2, 3] == 12.5
    assert np.isnan(gpp.neighbourhood_quantile_fast(np.full([50, 50], np.nan), 0.5, 1, thr)).all()
    assert (gpp.neighbourhood_quantile_fast(np.zeros([50, 50]), 0.5, 1, thr) == 0).all()
    thresholds = [0, 0.1, 0.2, 0.5, 1, 2, 5, 10, 20, 50, 100]
    for q in (0, 0.001, 0.999, 1):
        np.testing.assert_array_almost_equal(gpp.neighbourhood_quantile_fast(np.zeros([10, 10]), q, 5, thresholds), np.zeros([10, 10]))
    assert np.isnan(gpp.neighbourhood_quantile_fast(np.ones([5, 5]), np.nan, 1, [0, 1])).all()
    assert np.isnan(gpp.neighbourhood_quantile_fast(np.ones([5, 5]), 0.5, 1, [])).all()
    for q in (-0.1, 1.1):
        with pytest.raises(ValueError):
            gpp.neighbourhood_quantile_fast(np.ones([5, 5]), q, 1, [0, 1])
    q2 = np.full([5, 5], 0.5)
    q2[1, 1] = 1.5
    with pytest.raises(ValueError):
        gpp.neighbourhood_quantile_fast(np.ones([5, 5]), q2, 1, [0, 1])


def test_quantile_fast_random_vs_oracle(gpp, orc):
    rng = np.random.default_rng(4)
    f = rng.gamma(0.5, 2.0, size=(150, 333)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.3] = 0          # many exact zeros: plateaus in the CDF
    f[rng.uniform(size=f.shape) < 0.02] = np.nan
    for thr in (np.linspace(0, 5, 20), [0, 0, 1, 1, 2], [3, 1, 2], [0.5], np.linspace(0, 8, 64)):
        for hw in (0, 2, 15):
            for q in (0.0, 0.37, 1.0):
                assert_bit_exact(gpp.neighbourhood_quantile_fast(f, q, hw, thr), orc.neighbourhood_quantile_fast(f, q, hw, thr),
                                 "thr=%s hw=%d q=%g" % (np.asarray(thr)[:3], hw, q))


def test_quantile_fast_packed_counter_path(gpp, orc):
    """Shapes / thresholds the packed-counter kernel takes (row length a multiple of 4 and >= 256, ascending
    thresholds, half-width <= 15): plateaus (exact zeros, duplicate thresholds), missing and infinite values,
    clipped windows, every quantile special case, a quantile field, and row tiles with halo. Bit-exact."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(5)
    for shape in ((150, 512), (70, 260), (90, 333), (45, 70)):   # the last two: plain-load form (odd / narrow rows)
        f = rng.gamma(0.5, 2.0, size=shape).astype(f32)
        f[rng.uniform(size=shape) < 0.3] = 0
        f[rng.uniform(size=shape) < 0.02] = np.nan
        f[rng.uniform(size=shape) < 0.002] = np.inf
        f[10:40, 100:140] = np.nan                      # windows without any valid value
        f[50:60, 200:230] = 1.0                         # constant patch: F jumps from 0 to 1
        for thr in (np.linspace(0, 5, 20), [0, 0, 1, 1, 2], [0.5], np.linspace(0, 8, 31), [0, 1], np.linspace(-1, 3, 11)):
            for hw in (1, 2, 7, 15):
                for q in (0.0, 0.25, 0.37, 0.5, 1.0):
                    got = gpp.neighbourhood_quantile_fast(f, q, hw, thr)
                    want = orc.neighbourhood_quantile_fast(f, q, hw, thr)
                    assert_bit_exact(got, want, "%s thr=%s.. (%d) hw=%d q=%g" % (shape, np.asarray(thr)[:2], len(thr), hw, q))
    # even window counts make F hit the quantile exactly (the plateau rules of gridpp::interpolate)
    f = (rng.uniform(size=(64, 256)) < 0.5).astype(f32)
    for hw in (1, 2, 3):
        for q in (0.0, 0.25, 0.5, 0.75, 1.0):
            for thr in ([0, 1], [0, 0.5, 1], [-1, 0, 0, 1, 1, 2], [0.5, 0.6, 0.7]):
                assert_bit_exact(gpp.neighbourhood_quantile_fast(f, q, hw, thr), orc.neighbourhood_quantile_fast(f, q, hw, thr),
                                 "binary field thr=%s hw=%d q=%g" % (thr, hw, q))
    # spatially varying quantile (with missing entries)
    f = rng.gamma(0.5, 2.0, size=(90, 300)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.05] = np.nan
    qf = rng.uniform(size=f.shape).astype(f32)
    qf[rng.uniform(size=f.shape) < 0.05] = np.nan
    qf[0, :10] = [0, 1, 0, 1, 0.5, 0.5, 0, 1, 0, 1]
    thr = np.linspace(0, 5, 12).astype(f32)
    assert_bit_exact(gpp.neighbourhood_quantile_fast(f, qf, 4, thr), orc.neighbourhood_quantile_fast(f, qf, 4, thr), "quantile field")
    # row tiles with halo equal the whole field
    ny, nx, hw = 200, 260, 7
    f = rng.gamma(0.5, 2.0, size=(ny, nx)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.02] = np.nan
    d = torch.from_numpy(f).cuda()
    thr = np.linspace(0, 5, 20).astype(f32)
    whole = gd.neighbourhood_quantile_fast(d, 0.5, hw, thr)
    tiles = []
    for r0, r1 in ((0, 64), (64, 65), (65, 200)):
        lo, hi = max(0, r0 - hw), min(ny, r1 + hw)
        tiles.append(gd.neighbourhood_quantile_fast(d[lo:hi].contiguous(), 0.5, hw, thr, row0=r0 - lo, n_rows_out=r1 - r0))
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(torch.cat(tiles), nan=-777.0), torch.nan_to_num(whole, nan=-777.0))
    assert_bit_exact(whole.cpu().numpy(), orc.neighbourhood_quantile_fast(f, 0.5, hw, thr), "tiles vs oracle")


def test_ensemble_forms(gpp, orc):
    """gridpp::neighbourhood(vec3, ...) (neighbourhood.cpp:12-27) and neighbourhood_quantile_fast(vec3, ...) (:411-527)
    against the golden fixture of the compiled reference and against the oracle on shapes that take the TMA kernels."""
    g = golden("ensemble_forms")
    f = g["field"]
    for hw in (0, 1, 4, 9):
        for name, st in STATS.items():
            got, want = gpp.neighbourhood(f, hw, st), g["nbh_hw%d__%s" % (hw, name)]
            if name in ("count", "min", "max"):
                assert_bit_exact(got, want, "ens nbh hw=%d %s" % (hw, name))
            else:
                assert_close(got, want, 4.0 * (1 if name == "mean" else (2 * hw + 1) ** 2), 1e-6, "ens nbh hw=%d %s" % (hw, name))
    for hw in (0, 2, 6):
        for q in (0.0, 0.3, 0.5, 0.9, 1.0):
            got, want = gpp.neighbourhood_quantile_fast(f, q, hw, g["thresholds"]), g["qf_hw%d__q%g" % (hw, q)]
            assert_close(got, want, 1.0, 1e-5, "ens qfast hw=%d q=%g" % (hw, q), allow_outliers=2)
    got = gpp.neighbourhood_quantile_fast(f, g["qfield"], 3, g["thresholds"])
    assert_close(got, g["qf_hw3__qfield"], 1.0, 1e-5, "ens qfast field", allow_outliers=2)
    # larger, TMA-eligible shape against the oracle
    rng = np.random.default_rng(11)
    f = (rng.gamma(0.5, 2.0, size=(120, 260, 1)) + rng.normal(size=(120, 260, 6)) * 0.3).astype(f32)
    f[rng.uniform(size=f.shape) < 0.03] = np.nan
    for st_name, st in (("mean", gpp.Mean), ("max", gpp.Max)):
        got, want = gpp.neighbourhood(f, 7, st), orc.neighbourhood_ens(f, 7, st)
        if st_name == "max":
            assert_bit_exact(got, want, "ens max 260")
        else:
            assert_close(got, want, 4.0, 1e-6, "ens mean 260")
    thr = np.linspace(0, 6, 13).astype(f32)
    got, want = gpp.neighbourhood_quantile_fast(f, 0.5, 7, thr), orc.neighbourhood_quantile_fast_ens(f, 0.5, 7, thr)
    assert_close(got, want, 1.0, 1e-5, "ens qfast 260", allow_outliers=3)
    # identical members reproduce the 2-D result (tests/test_neighbourhood.py:133-144 of the reference)
    v = rng.uniform(size=(80, 64)).astype(f32)
    v3 = np.repeat(v[:, :, None], 5, axis=2)
    np.testing.assert_array_almost_equal(gpp.neighbourhood(v, 5, gpp.Mean), gpp.neighbourhood(v3, 5, gpp.Mean), 5)
    np.testing.assert_array_almost_equal(gpp.neighbourhood_quantile_fast(v, 0.5, 5, [0, 0.25, 0.5, 0.75, 1]),
                                         gpp.neighbourhood_quantile_fast(v3, 0.5, 5, [0, 0.25, 0.5, 0.75, 1]))
    assert gpp.neighbourhood(np.zeros((0, 0, 3), f32), 1, gpp.Mean).shape == (0, 0)
    with pytest.raises(ValueError):
        gpp.neighbourhood_quantile_fast(v3, 1.5, 1, [0, 1])


def test_get_neighbourhood_thresholds(gpp, orc):
    """gridpp::get_neighbourhood_thresholds, neighbourhood.cpp:243-295 / calc_even_quantiles util.cpp:261-338: the
    reference's known answers (tests/test_get_neighbourhood_thresholds.py:8-51) and the oracle on random fields with
    missing values, long runs of the lowest value, more thresholds than values, and the 3-D form. Bit-exact."""
    for num in (-1, 0):
        with pytest.raises(ValueError):
            gpp.get_neighbourhood_thresholds(np.ones([5, 5]), num)
    for num in (1, 5):
        assert gpp.get_neighbourhood_thresholds(np.zeros((0, 0)), num).size == 0
    field = np.reshape(np.arange(4), [2, 2])
    for num in (4, 5, 6):
        np.testing.assert_array_equal(gpp.get_neighbourhood_thresholds(field, num), [0, 1, 2, 3])
    field = np.reshape([0, 0, 2, 3, 4, 5, 6, 11, 8, 9, 10, 11], [3, 4]).astype(f32)
    for num in range(1, 6):
        assert_bit_exact(gpp.get_neighbourhood_thresholds(field, num), orc.get_neighbourhood_thresholds(field, num), "duplicates num=%d" % num)
    rng = np.random.default_rng(21)
    f = rng.gamma(0.5, 2.0, size=(300, 257)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.05] = np.nan
    f[3, 3] = np.inf
    g = f.copy()
    g[rng.uniform(size=f.shape) < 0.4] = 0          # a long run of the lowest value (util.cpp:302-308)
    for name, fld in (("gamma", f), ("zeros", g), ("constant", np.full((20, 30), 2.5, f32)), ("all missing", np.full((4, 4), np.nan, f32))):
        for num in (1, 2, 3, 11, 20, 64):
            assert_bit_exact(gpp.get_neighbourhood_thresholds(fld, num), orc.get_neighbourhood_thresholds(fld, num), "%s num=%d" % (name, num))
    v = rng.uniform(size=(10, 10)).astype(f32)
    v3 = np.repeat(v[:, :, None], 5, axis=2)
    for num in (1, 5):
        np.testing.assert_array_almost_equal(gpp.get_neighbourhood_thresholds(v, num), gpp.get_neighbourhood_thresholds(v3, num))
    # thresholds from the device feed quantile_fast exactly like the oracle's
    thr = gpp.get_neighbourhood_thresholds(f, 11)
    assert_bit_exact(gpp.neighbourhood_quantile_fast(f, 0.9, 2, thr), orc.neighbourhood_quantile_fast(f, 0.9, 2, orc.get_neighbourhood_thresholds(f, 11)),
                     "quantile_fast on device thresholds")


# ------------------------------------------------------------------ full BASELINE.json sizes ------------
def test_full_size_neighbourhood_properties(gpp, orc):
    """Config 2 (4000 x 4000, halfwidth 7): the stencil is local, so any window of the full-size result must equal
    the oracle run on that window plus its halo; plus idempotence / ordering properties over the whole field."""
    rng = np.random.default_rng(1000)
    n = 4000
    f = (rng.random((n, n), dtype=f32) * 10).astype(f32)
    f[rng.random((n, n), dtype=f32) < 0.01] = np.nan
    f[1000:1020, 2000:2020] = np.nan
    hw = 7
    res = {name: gpp.neighbourhood(f, hw, st) for name, st in STATS.items()}
    for (r, c) in ((0, 0), (0, n - 90), (n - 90, 0), (n - 90, n - 90), (990, 1990), (2000, 1234)):
        r0, r1, c0, c1 = max(0, r - hw), min(n, r + 90 + hw), max(0, c - hw), min(n, c + 90 + hw)
        for name, st in STATS.items():
            want = orc.neighbourhood(f[r0:r1, c0:c1], hw, st)[r - r0:r - r0 + 90, c - c0:c - c0 + 90]
            got = res[name][r:r + 90, c:c + 90]
            if name in ("count", "min", "max"):
                assert_bit_exact(got, want, "window (%d,%d) %s" % (r, c, name))
            else:
                assert_close(got, want, 10.0 * (1 if name == "mean" else 225), 1e-6, "window (%d,%d) %s" % (r, c, name))
    ok = ~np.isnan(res["mean"])
    assert (res["min"][ok] <= res["mean"][ok] + 1e-4).all() and (res["mean"][ok] <= res["max"][ok] + 1e-4).all()
    assert np.isnan(res["mean"][1007:1013, 2007:2013]).all() and (res["count"][1007:1013, 2007:2013] == 0).all()
    assert res["count"].max() == 225 and res["count"][0, 0] <= 64
    assert_bit_exact(gpp.neighbourhood(res["max"], 0, gpp.Max), res["max"])


def test_full_size_oi_subsample_equals_oracle(gpp, orc):
    """Config 3 at full size (4000 x 4000 grid, dx 250 m, 10 000 observations, Barnes 10 km, max_points 30): every
    grid point is independent, so a random subsample of the full-grid analysis must match the oracle evaluated at
    just those points (the Points overload, oi.cpp:138)."""
    rng = np.random.default_rng(1000)
    n, dx, S = 4000, 250.0, 10000
    y, x = np.meshgrid(np.arange(n, dtype=f32) * dx, np.arange(n, dtype=f32) * dx, indexing="ij")
    py, px = (rng.random(S) * n * dx).astype(f32), (rng.random(S) * n * dx).astype(f32)
    bg = rng.standard_normal((n, n), dtype=f32) * 3
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    pbg = gpp.nearest(grid, points, bg)
    obs = (pbg + rng.standard_normal(S, dtype=f32) * 0.5).astype(f32)
    ratios = np.full(S, 0.5, f32)
    out = gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, gpp.BarnesStructure(10000), 30)
    assert out.shape == (n, n) and not np.isnan(out).any()
    pick = rng.choice(n * n, 6000, replace=False)
    pick[:4] = [0, n - 1, n * (n - 1), n * n - 1]
    want = orc.optimal_interpolation((y.ravel()[pick], x.ravel()[pick], None, None), bg.ravel()[pick], (py, px, None, None), obs, ratios,
                                     pbg, B.make_structure(B.BARNES, 10000.0), 30, B.CARTESIAN)
    assert_close(out.ravel()[pick], want, 3.0, RTOL, "C3 subsample")
    assert np.abs(out - bg).max() < 10.0
    # pbackground from the device nearest() equals the oracle's nearest neighbour on a subsample of the observations
    sub = slice(0, 300)
    idx = orc.points_nearest(y[:400, :], x[:400, :], B.CARTESIAN, np.minimum(py[sub], 399 * dx), px[sub])
    got = grid._set.nearest(np.minimum(py[sub], 399 * dx), px[sub])
    assert_bit_exact(got, idx, "nearest on the 16M-node grid")


# ------------------------------------------------------------------ ensemble OI (EnSI) ------------------
def test_ensi_golden(gpp):
    g = golden("ensi_c5_density")
    grid = gpp.Grid(g["y"], g["x"], type=gpp.Cartesian)
    points = gpp.Points(g["py"], g["px"], type=gpp.Cartesian)
    s = product_structure(gpp, parse_spec(g["structure"]))
    scale = float(np.std(g["background"]))
    for name in ("mp20", "mp20_clamp", "unlimited"):
        mp, extr = (int(v) for v in g[name + "__args"])
        out = gpp.optimal_interpolation_ensi(grid, g["background"], points, g["pobs"], g["psigmas"], g["pbackground"], s, mp, bool(extr))
        assert out.shape == g["background"].shape and out.dtype == np.float32
        assert_close(out, g[name + "__analysis"], scale, RTOL, "EnSI " + name)


def test_ensi_known_answers(gpp):
    # tests/test_optimal_interpolation_ens.py:9-35 of the reference
    rng = np.random.default_rng(2)
    y, x = np.meshgrid(np.arange(6) * 1000.0, np.arange(7) * 1000.0, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    bg = rng.normal(size=(6, 7, 5)).astype(f32)
    s = gpp.BarnesStructure(2500)
    out = gpp.optimal_interpolation_ensi(grid, bg, gpp.Points([], [], type=gpp.Cartesian), [], [], np.zeros((0, 5)), s, 10)
    assert_bit_exact(out, bg)
    points = gpp.Points([1000, 3000], [1000, 4000], type=gpp.Cartesian)
    pbg = rng.normal(size=(2, 5)).astype(f32)
    out = gpp.optimal_interpolation_ensi(grid, bg, points, [1.0, np.nan], [0.5, 0.5], pbg, s, 10)
    assert not np.isnan(out).any() and np.abs(out - bg).max() > 0
    one = gpp.optimal_interpolation_ensi(grid, bg, gpp.Points([1000], [1000], type=gpp.Cartesian), [1.0], [0.5], pbg[:1], s, 10)
    assert_bit_exact(out, one)
    # a member with an invalid value anywhere is left untouched (oi_ensi.cpp:187-201)
    bg2 = bg.copy()
    bg2[2, 3, 1] = np.nan
    out2 = gpp.optimal_interpolation_ensi(grid, bg2, points, [1.0, 0.5], [0.5, 0.5], pbg, s, 10)
    assert_bit_exact(out2[:, :, 1], bg2[:, :, 1])
    assert np.abs(out2[:, :, 0] - bg2[:, :, 0]).max() > 0
    with pytest.raises(ValueError):
        gpp.optimal_interpolation_ensi(grid, bg, points, [1.0, 0.5], [0.5, 0.5], pbg, s, -1)


def test_ensi_random_vs_oracle(gpp, orc):
    """BASELINE.json config 5 density (dx 200 m, 20 members, Barnes 10 km, max_points 50, ~84 candidates) on a sub-grid."""
    rng = np.random.default_rng(1000)
    ny, nx, dx, E = 40, 50, 200.0, 20
    y, x = np.meshgrid(30000 + np.arange(ny) * dx, 30000 + np.arange(nx) * dx, indexing="ij")
    S = 1400   # 0.02 obs / km^2 around the sub-grid
    py, px = rng.uniform(-10000, 80000, S).astype(f32), rng.uniform(-10000, 80000, S).astype(f32)
    bg = (rng.normal(size=(ny, nx, 1)) * 2 + rng.normal(size=(ny, nx, E))).astype(f32)
    pbg = rng.normal(size=(S, E)).astype(f32)
    obs = rng.normal(size=S).astype(f32)
    obs[rng.uniform(size=S) < 0.01] = np.nan
    sig = np.full(S, 0.5, f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    for mp, extr in ((50, True), (50, False), (7, True)):
        got = gpp.optimal_interpolation_ensi(grid, bg, points, obs, sig, pbg, gpp.BarnesStructure(10000), mp, extr)
        want = orc.optimal_interpolation_ensi((y, x, None, None), bg, (py, px, None, None), obs, sig, pbg,
                                              B.make_structure(B.BARNES, 10000.0), mp, B.CARTESIAN, allow_extrapolation=extr)
        assert_close(got.reshape(-1, E), want, 2.0, RTOL, "EnSI mp=%d extr=%s" % (mp, extr))
    # the other ensemble sizes with their own kernel instantiation (10, 30) and the largest generic one (31)
    for E2 in (10, 30, 31):
        sub = (slice(0, 12), slice(0, 15))
        bg2 = (rng.normal(size=(12, 15, 1)) * 2 + rng.normal(size=(12, 15, E2))).astype(f32)
        pbg2 = rng.normal(size=(S, E2)).astype(f32)
        got = gpp.optimal_interpolation_ensi(gpp.Grid(y[sub], x[sub], type=gpp.Cartesian), bg2, points, obs, sig, pbg2, gpp.BarnesStructure(10000), 40)
        want = orc.optimal_interpolation_ensi((y[sub], x[sub], None, None), bg2, (py, px, None, None), obs, sig, pbg2,
                                              B.make_structure(B.BARNES, 10000.0), 40, B.CARTESIAN)
        assert_close(got.reshape(-1, E2), want, 2.0, RTOL, "EnSI %d members" % E2)
    # Points overload, 7 members (odd: exercises the padded rotation schedule), Geodetic
    la, lo = rng.uniform(59, 60, 300).astype(f32), rng.uniform(10, 12, 300).astype(f32)
    pla, plo = rng.uniform(59, 60, 150).astype(f32), rng.uniform(10, 12, 150).astype(f32)
    bgp, pbgp = rng.normal(size=(300, 7)).astype(f32), rng.normal(size=(150, 7)).astype(f32)
    o, sg = rng.normal(size=150).astype(f32), rng.uniform(0.3, 1.0, 150).astype(f32)
    got = gpp.optimal_interpolation_ensi(gpp.Points(la, lo), bgp, gpp.Points(pla, plo), o, sg, pbgp, gpp.BarnesStructure(20000), 12)
    want = orc.optimal_interpolation_ensi((la, lo, None, None), bgp, (pla, plo, None, None), o, sg, pbgp,
                                          B.make_structure(B.BARNES, 20000.0), 12, B.GEODETIC)
    assert_close(got, want, 1.0, RTOL, "EnSI points geodetic")


def test_ensi_pipelined_path_equals_whole_field_path(gpp, orc):
    """Fields of 2^18 points or more go through the pipelined host path (member scan on the host, per-block upload,
    overlapping block kernels, staged download); every point is independent, so its result must equal what the
    small-field path gives on the two halves of the grid, and the oracle on a sample. One member carries a NaN and
    must come back untouched (oi_ensi.cpp:187-201)."""
    rng = np.random.default_rng(11)
    ny, nx, dx, E, S = 520, 520, 500.0, 6, 400
    assert ny * nx >= 1 << 18 and ny * nx // 2 < 1 << 18
    y, x = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    y, x = y.astype(f32), x.astype(f32)
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    bg = (rng.normal(size=(ny, nx, 1)) + rng.normal(size=(ny, nx, E))).astype(f32)
    bg[400, 17, 3] = np.nan
    pbg = rng.normal(size=(S, E)).astype(f32)
    obs, sig = rng.normal(size=S).astype(f32), np.full(S, 0.6, f32)
    points, s = gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(12000)
    whole = gpp.optimal_interpolation_ensi(gpp.Grid(y, x, type=gpp.Cartesian), bg, points, obs, sig, pbg, s, 12)
    assert_bit_exact(whole[:, :, 3], bg[:, :, 3], "member with a missing value is left untouched")
    h = ny // 2
    # the lower half holds the missing value, so it also runs with five valid members
    lower = gpp.optimal_interpolation_ensi(gpp.Grid(y[h:], x[h:], type=gpp.Cartesian), bg[h:], points, obs, sig, pbg, s, 12)
    keep = [e for e in range(E) if e != 3]
    assert_close(whole[h:][:, :, keep], lower[:, :, keep], 1.0, 1e-6, "pipelined vs whole-field path")
    pick = rng.choice(h * nx, 300, replace=False)
    pick[0] = (400 - h) * nx + 17   # the sample must see the missing value too, or it would analyse six members
    want = orc.optimal_interpolation_ensi((y[h:].ravel()[pick], x[h:].ravel()[pick], None, None), bg[h:].reshape(-1, E)[pick], (py, px, None, None),
                                          obs, sig, pbg, B.make_structure(B.BARNES, 12000.0), 12, B.CARTESIAN)
    assert_close(whole[h:].reshape(-1, E)[pick][:, keep], want[:, keep], 1.0, RTOL, "pipelined path vs oracle")


def test_structure_functions_with_point_objects(gpp):
    """tests/test_barnes_structure.py:8-33 and :46-66 of the reference as written there: gridpp.Point arguments, scalar
    results, the vector overload, and corr() of a spatially varying structure function (scales of the node nearest to p1)."""
    x = [0, 1000, 2000, 3000, np.nan]
    barnes = gpp.BarnesStructure(2000)
    cases = [(barnes, [1, 0.8824968934059143, 0.6065306663513184, 0.32465246319770813, 0], False),
             (gpp.CressmanStructure(2000), [1, 0.6, 0, 0, 0], False),
             (gpp.CrossValidation(barnes, 1000), [0, 0, 0.6065306663513184, 0.32465246319770813, 0], True)]
    for s, want, is_cv in cases:
        for xi, w in zip(x, want):
            p1 = gpp.Point(0, 0, 0, 0, gpp.Cartesian)
            p2 = gpp.Point(xi, 0, 0, 0, gpp.Cartesian)
            for func in ([s.corr_background] if is_cv else [s.corr, s.corr_background]):
                assert abs(func(p1, p2) - w) < 1e-7, (type(s).__name__, xi)
                assert abs(func(p2, p1) - w) < 1e-7, (type(s).__name__, xi)
                if not is_cv and not np.isnan(xi):
                    assert abs(func(p2, p2) - 1) < 1e-7
    others = [gpp.Point(v, 0, 0, 0, gpp.Cartesian) for v in x[:4]]
    np.testing.assert_array_equal(barnes.corr(gpp.Point(0, 0, 0, 0, gpp.Cartesian), others),
                                  np.array([1, 0.8824968934059143, 0.6065306663513184, 0.32465246319770813], f32))
    grid = gpp.Grid([[0, 0]], [[0, 2500]], [[0, 0]], [[0, 0]], gpp.Cartesian)
    s = gpp.BarnesStructure(grid, [[2500, 1]], [[0, 0]], [[0, 0]], 0.1)
    p1, p2 = gpp.Point(0, 0, 0, 0, gpp.Cartesian), gpp.Point(0, 2500, 0, 0, gpp.Cartesian)
    assert abs(s.localization_distance(p1) - np.sqrt(-2 * np.log(0.1)) * 2500) < 1e-2      # the scale at p1 is 2500
    assert abs(s.corr(p1, p2) - 0.6065306663513184) < 1e-6
    assert abs(s.localization_distance(p2) - np.sqrt(-2 * np.log(0.1)) * 1) < 1e-4         # the scale at p2 is 1
    assert s.corr(p2, p1) == 0
