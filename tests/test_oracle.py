"""Pins oracle/gridpp_oracle.c (the plain-C restatement the GPU parity tests check against).

Three anchors, none of which needs a GPU:
  1. known-answer values held by the reference's own tests (cited per test as tests/<file>:<line> of metno/gridpp);
  2. tests/golden/*.npz, generated from the reference sources themselves by tests/golden/make_golden.py;
  3. the compiled reference (oracle/_ref/libgridpp_ref.so) on fresh random inputs, when it is present.
"""
import numpy as np
import pytest

from oracle import bindings as B
from util import assert_bit_exact, assert_close, golden, oracle_structure, parse_spec

f32 = np.float32


# ------------------------------------------------------------------ 1. reference known answers ---------
def test_oi_single_observation(orc):
    # tests/test_optimal_interpolation.py:50-63 and :65-105
    y, x = [0, 0, 0], [0, 2500, 10000]
    s = B.make_structure(B.BARNES, 2500)
    out, var = orc.optimal_interpolation((y, x, y, y), np.zeros(3), ([0], [2500], [0], [0]), [1], [0.1], [0], s, 10,
                                         B.CARTESIAN, want_variance=True)
    np.testing.assert_array_almost_equal(out, [np.exp(-0.5) / 1.1, 1 / 1.1, np.exp(-0.5 * 9) / 1.1])
    assert abs(var[1] - 0.1 / 1.1) < 1e-7


def test_oi_no_observations_returns_background(orc):
    # tests/test_optimal_interpolation.py:192-202
    bg = np.arange(6, dtype=f32)
    out = orc.optimal_interpolation((np.zeros(6), np.arange(6) * 1000.0, None, None), bg, ([], [], None, None), [], [], [],
                                    B.make_structure(B.BARNES, 2500), 10, B.CARTESIAN)
    assert_bit_exact(out, bg)


def test_oi_nan_observation_ignored(orc):
    # tests/test_optimal_interpolation.py:154-168
    y, x = np.zeros(4), np.arange(4) * 1000.0
    s = B.make_structure(B.BARNES, 2500)
    args = dict(structure=s, max_points=10, ctype=B.CARTESIAN)
    a = orc.optimal_interpolation((y, x, None, None), np.zeros(4), (y[:2], x[:2], None, None), [1, np.nan], [0.5, 0.5], [0, 0], **args)
    b = orc.optimal_interpolation((y, x, None, None), np.zeros(4), (y[:1], x[:1], None, None), [1], [0.5], [0], **args)
    assert_bit_exact(a, b)


def test_barnes_goldens(orc):
    # tests/test_barnes_structure.py:8-33 (bit-exact float32 values) and :85-97 (hmax truncation)
    x = [0, 1000, 2000, 3000, np.nan]
    want = {"barnes": [1, 0.8824968934059143, 0.6065306663513184, 0.32465246319770813, 0],
            "cressman": [1, 0.6, 0, 0, 0], "cv": [0, 0, 0.6065306663513184, 0.32465246319770813, 0]}
    barnes = B.make_structure(B.BARNES, 2000)
    structs = {"barnes": barnes, "cressman": B.make_structure(B.CRESSMAN, 2000), "cv": B.cross_validation(barnes, 1000)}
    p1 = np.zeros((5, 5), f32)
    p2 = np.zeros((5, 5), f32)
    p2[:, 1] = x     # Point(lat=x, lon=0, Cartesian): y = lat
    for name, s in structs.items():
        got = orc.structure_corr(s, p1, p2, background=True)
        np.testing.assert_array_equal(got, np.array(want[name], f32), err_msg=name)
        np.testing.assert_array_equal(orc.structure_corr(s, p2, p1, background=True), np.array(want[name], f32))
    ans = {0: 1, 1000: 0.8824968934059143, 2000: 0.6065306663513184, 3000: 0.32465246319770813}
    for hmax in (0, 1000, 2000, 10000):
        s = B.make_structure(B.BARNES, 2000, 0, 0, hmax)
        for dist, a in ans.items():
            got = orc.structure_corr(s, [[0, 0, 0, 0, 0]], [[0, dist, 0, 0, 0]])[0]
            assert got == (f32(0) if dist > hmax else f32(a)), (hmax, dist, got)


def test_structure_invalid_elevation_ignored(orc):
    # tests/test_structure.py:24-36
    for st in (B.BARNES, B.CRESSMAN):
        for v in (0, 100):
            s = B.make_structure(st, 2000, v)
            a = orc.structure_corr(s, [[0, 0, 0, 0, 0]], [[0, 1000, 0, 0, 0]])
            b = orc.structure_corr(s, [[0, 0, 0, 0, 0]], [[0, 1000, 0, np.nan, 0]])
            assert a[0] == b[0]


def test_multiple_structure(orc):
    # tests/test_structure.py:91-120: Cressman(2000,...) x Cressman(200,...) x Cressman(2,...)
    s = B.multiple_structure(B.make_structure(B.CRESSMAN, 2000, 2000, 2000), B.make_structure(B.CRESSMAN, 200, 200, 200),
                             B.make_structure(B.CRESSMAN, 2, 2, 2))
    p0 = [0, 0, 0, 0, 0]
    assert abs(orc.structure_corr(s, [p0], [[0, 1000, 0, 0, 0]])[0] - 0.6) < 1e-6     # horizontal only
    assert abs(orc.structure_corr(s, [p0], [[0, 0, 0, 100, 0]])[0] - 0.6) < 1e-6      # vertical only (elev 0 -> -100)
    assert abs(orc.structure_corr(s, [p0], [[0, 0, 0, 0, 1]])[0] - 0.6) < 1e-6        # laf only


def test_structure_constructor_validation(orc):
    # tests/test_structure.py:8-62, tests/test_barnes_structure.py:35-44
    for st in (B.BARNES, B.CRESSMAN):
        for bad in (-1, np.nan):
            with pytest.raises(ValueError):
                orc.structure_describe(st, bad)
            with pytest.raises(ValueError):
                orc.structure_describe(st, 2000, bad)
            with pytest.raises(ValueError):
                orc.structure_describe(st, 2000, 100, bad)
    with pytest.raises(ValueError):
        orc.structure_describe(B.BARNES, 2000, 100, 0, -1)


def test_kdtree_known_answers(orc):
    # tests/test_kdtree.py:8-13
    idx, _, cnt = orc.points_neighbours([60, 61, 62], [10, 10, 12], B.GEODETIC, [60, 60], [10, 10], [1, 112000], capacity=4)
    assert cnt.tolist() == [1, 2] and idx[0, 0] == 0 and idx[1, :2].tolist() == [0, 1]
    # :15-18 distance golden
    _, dist, _ = orc.points_neighbours([0, 1000, 2000], [0, 1000, 2000], B.CARTESIAN, [100], [100], [1000], capacity=4)
    assert dist[0, 0] == f32(100 * np.sqrt(2))
    # :21-34 duplicates
    _, _, cnt = orc.points_neighbours([50, 50, 51], [0, 0, 10], B.GEODETIC, [50, 50], [0.001, 0], [1000, 1000], capacity=4)
    assert cnt.tolist() == [2, 2]
    # :35-53 poles
    idx, dist, cnt = orc.points_neighbours([89, 89, 90, 90], [0, 180, 0, 10], B.GEODETIC, [90], [0], [1000], capacity=4)
    assert cnt[0] == 2 and sorted(idx[0, :2].tolist()) == [2, 3] and abs(dist[0, 0]) < 1e-3 and abs(dist[0, 1]) < 1e-3
    # :149-161 radius edge: strictly inside the box AND distance <= radius
    la, lo = [0, 1000, 2000], [0, 0, 0]
    q = lambda lat, r, m=True: orc.points_neighbours_raw(la, lo, B.CARTESIAN, lat, 0, r, m, 8).tolist()
    assert q(900, 501) == [1] and q(900, 99.99) == [] and q(0, 1000) == [0] and q(0, 1001) == [0, 1] and q(0, 1001, False) == [1]
    # :108-112
    assert abs(orc.calc_distance(0, 0, 0.001, 0.001, B.GEODETIC) - 157.42953491210938) < 1e-4
    x, y, z = orc.convert_coordinates([0, 0.001], [0, 0.001], B.GEODETIC)
    d = np.sqrt(f32(x[0] - x[1]) ** 2 + f32(y[0] - y[1]) ** 2 + f32(z[0] - z[1]) ** 2)
    assert abs(d - 157.42953491210938) < 1e-4
    # :167-175 invalid coordinates
    for lat, lon in ((91, 0), (-91, 0), (np.nan, 0), (0, np.nan)):
        with pytest.raises(ValueError):
            orc.convert_coordinates([lat], [lon], B.GEODETIC)
    # :188-203 longitude wrap
    for lon in (-360, 0, 360):
        idx, dist, cnt = orc.points_neighbours([0], [lon], B.GEODETIC, [0, 0], [0, 180], [1e9, 1e9], capacity=2)
        assert cnt.tolist() == [1, 1] and abs(dist[0, 0]) < 1 and abs(dist[1, 0] - 12756274.0) < 2


def test_nearest_empty_and_out_of_domain(orc):
    # tests/test_points.py (get_nearest_neighbour of an empty set -> -1), tests/test_nearest.py (empty input -> NaN)
    assert orc.points_nearest([], [], B.GEODETIC, [0], [0]).tolist() == [-1]
    assert np.isnan(orc.nearest([], [], B.GEODETIC, [0, 1], [0, 1], np.zeros(0, f32))).all()
    la, lo = np.meshgrid([0, 1, 2], [0, 1, 2], indexing="ij")
    assert orc.points_nearest(la, lo, B.GEODETIC, [10, -10], [0.9, 2.4]).tolist() == [7, 2]


values5 = np.reshape(np.arange(25), [5, 5]).astype(f32)
values5[1, 3] = np.nan
values5[2, 4] = np.nan


def test_neighbourhood_known_answers(orc):
    # tests/test_neighbourhood.py:75-120
    m = orc.neighbourhood(values5, 1, B.MEAN)
    assert m[2, 2] == 12.5 and abs(m[0, 4] - 5.3333) < 1e-4
    assert (np.abs(orc.neighbourhood(values5, 100, B.MEAN) - 12.086956) < 1e-4).all()
    m0 = orc.neighbourhood(values5, 0, B.MEAN)
    assert_bit_exact(m0, values5)
    c = orc.neighbourhood(values5, 1, B.COUNT)
    assert c[2, 2] == 8 and c[0, 4] == 3 and (orc.neighbourhood(values5, 100, B.COUNT) == 23).all()
    mn, mx = orc.neighbourhood(values5, 1, B.MIN), orc.neighbourhood(values5, 1, B.MAX)
    assert mn[2, 2] == 6 and mn[0, 4] == 3 and mx[2, 2] == 18 and mx[0, 4] == 9
    assert (orc.neighbourhood(values5, 100, B.MIN) == 0).all() and (orc.neighbourhood(values5, 100, B.MAX) == 24).all()
    # :48-58 all-missing windows
    empty = np.zeros([5, 5], f32)
    empty[0:3, 0:3] = np.nan
    for st in (B.MEAN, B.MIN, B.MAX, B.SUM):
        assert np.isnan(orc.neighbourhood(empty, 1, st)[0:2, 0:2]).all()
    np.testing.assert_array_equal(orc.neighbourhood(empty, 1, B.COUNT),
                                  [[0, 0, 2, 4, 4], [0, 0, 3, 6, 6], [2, 3, 5, 7, 6], [4, 6, 7, 8, 6], [4, 6, 6, 6, 4]])
    # :146-152 no overflow for large values
    big = (np.arange(1, 1000, dtype=np.float64) ** 3).astype(f32)[:, None]
    np.testing.assert_array_almost_equal(orc.neighbourhood(big, 0, B.MEAN) / big - 1, np.zeros(big.shape), 6)
    with pytest.raises(ValueError):
        orc.neighbourhood(values5, -1, B.MEAN)
    with pytest.raises(ValueError):
        orc.neighbourhood(values5, 1, 40)


def test_neighbourhood_fast_equals_brute_force(orc):
    # the reference's own cross-check, tests/test_neighbourhood.py (every statistic, fast == brute force)
    rng = np.random.default_rng(1000)
    f = rng.uniform(size=(40, 33)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.05] = np.nan
    for hw in (0, 1, 3, 50):
        for st in (B.MIN, B.MAX, B.COUNT):
            assert_bit_exact(orc.neighbourhood(f, hw, st), orc.neighbourhood_brute_force(f, hw, st))
        assert_close(orc.neighbourhood(f, hw, B.MEAN), orc.neighbourhood_brute_force(f, hw, B.MEAN), 1.0, 2e-6)


def test_quantile_fast_known_answers(orc):
    # tests/test_neighbourhood_quantile_fast.py:52-58, :66-85, :129-136
    field = np.reshape(np.arange(9), [3, 3]).astype(f32)
    for hw in (0, 1, 2):
        np.testing.assert_array_equal(orc.neighbourhood_quantile_fast(field, 0.9, hw, [0]), np.zeros([3, 3]))
    thr = orc.get_neighbourhood_thresholds(values5, 100)
    out = orc.neighbourhood_quantile_fast(values5, 0.5, 1, thr)
    assert out[2, 2] == 12 and out[2, 3] == 12.5          # the reference's documented quirk values
    assert np.isnan(orc.neighbourhood_quantile_fast(np.full([20, 20], np.nan, f32), 0.5, 1, thr)).all()
    assert (orc.neighbourhood_quantile_fast(np.zeros([20, 20], f32), 0.5, 1, thr) == 0).all()
    empty = np.zeros([5, 5], f32)
    empty[0:3, 0:3] = np.nan
    assert np.isnan(orc.neighbourhood_quantile_fast(empty, 0.5, 1, [0, 1])[0:2, 0:2]).all()
    thresholds = [0, 0.1, 0.2, 0.5, 1, 2, 5, 10, 20, 50, 100]
    for q in (0, 0.001, 0.999, 1):
        np.testing.assert_array_almost_equal(orc.neighbourhood_quantile_fast(np.zeros([10, 10], f32), q, 5, thresholds), np.zeros([10, 10]))
    assert np.isnan(orc.neighbourhood_quantile_fast(np.ones([5, 5], f32), np.nan, 1, [0, 1])).all()   # :34-41
    assert np.isnan(orc.neighbourhood_quantile_fast(np.ones([5, 5], f32), 0.5, 1, [])).all()
    for q in (-0.1, 1.1):
        with pytest.raises(ValueError):
            orc.neighbourhood_quantile_fast(np.ones([5, 5], f32), q, 1, [0, 1])
    with pytest.raises(ValueError):
        orc.neighbourhood_quantile_fast(np.ones([5, 5], f32), 0.5, -1, [0, 1])


def test_thresholds_and_interpolate_known_answers(orc):
    # tests/test_get_neighbourhood_thresholds.py:21-27
    np.testing.assert_array_equal(orc.get_neighbourhood_thresholds(np.array([[0, 1], [2, 3]], f32), 4), [0, 1, 2, 3])
    # tests/test_interpolate.py:8-67
    x, y = [0, 1, 2], [0, 2, 1]
    assert [orc.interpolate(v, x, y) for v in (0, 2, 1, -1, 3)] == [0, 1, 2, 0, 1]
    assert abs(orc.interpolate(0.9, x, y) - 1.8) < 1e-6 and abs(orc.interpolate(1.5, x, y) - 1.5) < 1e-6
    dx, dy = [0, 0, 0.5, 0.5, 1, 1], [0, 0.1, 0.4, 0.6, 0.9, 1]
    assert abs(orc.interpolate(1, dx, dy) - 0.9) < 1e-6 and abs(orc.interpolate(0, dx, dy) - 0.1) < 1e-6
    assert abs(orc.interpolate(0.5, dx, dy) - 0.5) < 1e-6
    assert abs(orc.interpolate(0.499999, dx, dy) - 0.4) < 1e-5 and abs(orc.interpolate(0.500001, dx, dy) - 0.6) < 1e-5
    assert abs(orc.interpolate(0, [0] * 6, dy) - 0.5) < 1e-6
    assert np.isnan(orc.interpolate(0, [], [])) and np.isnan(orc.interpolate(np.nan, [0], [0]))


# ------------------------------------------------------------------ 2. golden fixtures -----------------
def test_golden_structure(orc):
    g = golden("structure")
    for name in g["names"]:
        s = oracle_structure(parse_spec(g[name + "__structure"]))
        assert_bit_exact(orc.structure_corr(s, g["p1"], g["p2"], False), g[name + "__corr"], name)
        assert_bit_exact(orc.structure_corr(s, g["p1"], g["p2"], True), g[name + "__corr_background"], name)
        assert f32(orc.structure_localization_distance(s)) == g[name + "__loc_dist"], name


def test_golden_index_queries(orc):
    g = golden("index_queries")
    for tname, t in (("geodetic", B.GEODETIC), ("cartesian", B.CARTESIAN)):
        k = lambda n: g[tname + "__" + n]
        for got, want in zip(orc.convert_coordinates(k("lats"), k("lons"), t), (k("x"), k("y"), k("z"))):
            assert_bit_exact(got, want, tname + " coordinates")
        assert_bit_exact(orc.points_nearest(k("lats"), k("lons"), t, k("qlats"), k("qlons")), k("nearest"))
        assert_bit_exact(orc.points_nearest(k("lats"), k("lons"), t, k("lats")[:200], k("lons")[:200], False), k("nearest_nomatch"))
        idx, dist, cnt = orc.points_neighbours(k("lats"), k("lons"), t, k("qlats"), k("qlons"), float(k("radius")), capacity=96)
        assert_bit_exact(cnt, k("nbr_count"))
        assert_bit_exact(idx, k("nbr_index"))
        assert_bit_exact(dist, k("nbr_dist"))
        assert_bit_exact(orc.points_closest(k("lats"), k("lons"), t, k("qlats"), k("qlons"), 5), k("closest5"))
        assert_bit_exact(orc.nearest(k("lats"), k("lons"), t, k("qlats"), k("qlons"), k("values")), k("nearest_values"))


def test_golden_oi(orc):
    g = golden("oi_c1_geodetic")
    s = oracle_structure(parse_spec(g["structure"]))
    out, var = orc.optimal_interpolation((g["lats"], g["lons"], None, None), g["background"], (g["plats"], g["plons"], None, None),
                                         g["pobs"], g["pratios"], g["pbackground"], s, int(g["max_points"]), int(g["ctype"]),
                                         want_variance=True)
    assert_bit_exact(out.reshape(g["analysis"].shape), g["analysis"], "C1 analysis")
    assert_bit_exact(var.reshape(g["analysis"].shape), g["analysis_variance"], "C1 variance")
    g = golden("oi_c3_density")
    for name in g["names"]:
        s = oracle_structure(parse_spec(g[name + "__structure"]))
        mp, extr, use_elev = (int(v) for v in g[name + "__args"])
        be, bl, pe, pl = (g["belev"], g["blaf"], g["pelev"], g["plaf"]) if use_elev else (None,) * 4
        out, var = orc.optimal_interpolation((g["y"], g["x"], be, bl), g["background"], (g["py"], g["px"], pe, pl), g["pobs"],
                                             g["pratios"], g["pbackground"], s, mp, B.CARTESIAN, allow_extrapolation=bool(extr),
                                             want_variance=True)
        assert_bit_exact(out.reshape(g["background"].shape), g[name + "__analysis"], name)
        assert_bit_exact(var.reshape(g["background"].shape), g[name + "__variance"], name + " variance")


def test_golden_ensi(orc):
    g = golden("ensi_c5_density")
    s = oracle_structure(parse_spec(g["structure"]))
    for name in ("mp20", "mp20_clamp", "unlimited"):
        mp, extr = (int(v) for v in g[name + "__args"])
        out = orc.optimal_interpolation_ensi((g["y"], g["x"], None, None), g["background"], (g["py"], g["px"], None, None), g["pobs"],
                                             g["psigmas"], g["pbackground"], s, mp, B.CARTESIAN, allow_extrapolation=bool(extr))
        assert_bit_exact(out.reshape(g["background"].shape), g[name + "__analysis"], name)


def test_golden_neighbourhood(orc):
    g = golden("neighbourhood")
    stats = {"mean": B.MEAN, "sum": B.SUM, "count": B.COUNT, "min": B.MIN, "max": B.MAX}
    for hw in (0, 1, 7, 15, 70):
        for name, st in stats.items():
            assert_bit_exact(orc.neighbourhood(g["field"], hw, st), g["hw%d__%s" % (hw, name)], "hw=%d %s" % (hw, name))


def test_golden_quantile_fast(orc):
    g = golden("quantile_fast")
    for hw in (1, 7, 15):
        for q in (0.0, 0.001, 0.5, 0.9, 0.999, 1.0):
            assert_bit_exact(orc.neighbourhood_quantile_fast(g["field"], q, hw, g["thresholds"]), g["hw%d__q%g" % (hw, q)])
    assert_bit_exact(orc.neighbourhood_quantile_fast(g["field"], g["qfield"], 3, g["thresholds"]), g["hw3__qfield"])
    assert_bit_exact(orc.get_neighbourhood_thresholds(g["field"], 11), g["thresholds_auto"])
    assert_bit_exact(orc.neighbourhood_quantile_fast(g["field"], 0.9, 2, g["thresholds_auto"]), g["hw2__auto_q0.9"])


# ------------------------------------------------------------------ 3. live against the compiled reference
def test_golden_ensemble_forms(orc):
    """gridpp::neighbourhood(vec3, ...) neighbourhood.cpp:12-27 and neighbourhood_quantile_fast(vec3, ...) :411-527:
    the C restatement against fixtures generated from the compiled reference (tests/golden/make_golden_ensemble.py)."""
    g = golden("ensemble_forms")
    f = g["field"]
    for hw in (0, 1, 4, 9):
        for name, st in (("mean", B.MEAN), ("sum", B.SUM), ("count", B.COUNT), ("min", B.MIN), ("max", B.MAX)):
            assert_bit_exact(orc.neighbourhood_ens(f, hw, st), g["nbh_hw%d__%s" % (hw, name)], "ens nbh hw=%d %s" % (hw, name))
    for hw in (0, 2, 6):
        for q in (0.0, 0.3, 0.5, 0.9, 1.0):
            assert_bit_exact(orc.neighbourhood_quantile_fast_ens(f, q, hw, g["thresholds"]), g["qf_hw%d__q%g" % (hw, q)],
                             "ens qfast hw=%d q=%g" % (hw, q))
    assert_bit_exact(orc.neighbourhood_quantile_fast_ens(f, g["qfield"], 3, g["thresholds"]), g["qf_hw3__qfield"], "ens qfast field")
    # tests/test_neighbourhood.py:133-144 and tests/test_neighbourhood_quantile_fast.py:92-104 of the reference:
    # identical members reproduce the 2-D result
    rng = np.random.RandomState(1000)
    values = rng.rand(60, 50).astype(np.float32)
    values3 = np.repeat(values[:, :, None], 5, axis=2)
    for hw in (0, 1, 5):
        np.testing.assert_array_almost_equal(orc.neighbourhood(values, hw, B.MEAN), orc.neighbourhood_ens(values3, hw, B.MEAN), 5)
        np.testing.assert_array_almost_equal(orc.neighbourhood_quantile_fast(values, 0.5, hw, [0, 0.25, 0.5, 0.75, 1]),
                                             orc.neighbourhood_quantile_fast_ens(values3, 0.5, hw, [0, 0.25, 0.5, 0.75, 1]))


def test_live_against_reference(orc, ref):
    rng = np.random.default_rng(7)
    ny = nx = 30
    yy, xx = np.meshgrid(np.arange(ny) * 3000.0, np.arange(nx) * 3000.0, indexing="ij")
    S = 200
    py, px = rng.uniform(0, ny * 3000, S).astype(f32), rng.uniform(0, nx * 3000, S).astype(f32)
    bg = rng.normal(size=ny * nx).astype(f32)
    pbg = rng.normal(size=S).astype(f32)
    obs = (pbg + rng.normal(size=S)).astype(f32)
    ratios = rng.uniform(0.1, 2, S).astype(f32)
    for st, mp in ((B.make_structure(B.BARNES, 8000.0), 25), (B.make_structure(B.TOAR, 2500.0), 8)):
        a = ref.optimal_interpolation((yy, xx, None, None), bg, (py, px, None, None), obs, ratios, pbg, st, mp, B.CARTESIAN, want_variance=True)
        b = orc.optimal_interpolation((yy, xx, None, None), bg, (py, px, None, None), obs, ratios, pbg, st, mp, B.CARTESIAN, want_variance=True)
        assert_bit_exact(a[0], b[0])
        assert_bit_exact(a[1], b[1])
    E = 5
    bgE, pbgE = rng.normal(size=(ny * nx, E)).astype(f32), rng.normal(size=(S, E)).astype(f32)
    sig = np.full(S, 0.5, f32)
    for extr in (True, False):
        a = ref.optimal_interpolation_ensi((yy, xx, None, None), bgE, (py, px, None, None), obs, sig, pbgE, B.make_structure(B.BARNES, 8000.0), 12, B.CARTESIAN, allow_extrapolation=extr)
        b = orc.optimal_interpolation_ensi((yy, xx, None, None), bgE, (py, px, None, None), obs, sig, pbgE, B.make_structure(B.BARNES, 8000.0), 12, B.CARTESIAN, allow_extrapolation=extr)
        assert_bit_exact(a, b)
    f = rng.uniform(size=(70, 45)).astype(f32)
    f[rng.uniform(size=f.shape) < 0.03] = np.nan
    for hw in (0, 2, 9):
        for stt in (B.MEAN, B.SUM, B.COUNT, B.MIN, B.MAX):
            assert_bit_exact(ref.neighbourhood(f, hw, stt), orc.neighbourhood(f, hw, stt))
        assert_bit_exact(ref.neighbourhood_quantile_fast(f, 0.3, hw, np.linspace(0, 1, 7)), orc.neighbourhood_quantile_fast(f, 0.3, hw, np.linspace(0, 1, 7)))


def test_golden_spatially_varying_structures(orc):
    """<Family>Structure(Grid, h, v, w, min_rho) (structure.cpp:168-184, :342, :492, :643, :790) through
    optimal_interpolation_full: the C restatement against the fixtures generated from the compiled reference
    (tests/golden/make_golden_spatial.py). The metric is the one of the GPU test: with Soar / Toar and elevations the
    systems are nearly singular at a few grid points, so the scale is the field's largest magnitude and at most 3 of the
    1440 points may exceed 1e-5 (none 1e-3)."""
    g = golden("oi_spatial_structure")
    for case in g["cases"]:
        name, stype, elev, mp, extr, min_rho = str(case).split(",")
        stype, elev, mp, extr, min_rho = int(stype), int(elev), int(mp), int(extr), float(min_rho)
        bpts = (g["y"], g["x"], g["belev"] if elev else None, g["blaf"] if elev else None)
        opts = (g["py"], g["px"], g["pelev"] if elev else None, g["plaf"] if elev else None)
        out, var = orc.optimal_interpolation_spatial(bpts, g["background"], opts, g["pobs"], g["pratios"], g["pbackground"], stype,
                                                     (g["gy"], g["gx"]), g["h"], g["v"], g["w"], min_rho, mp, B.CARTESIAN,
                                                     allow_extrapolation=bool(extr), want_variance=True)
        key = "%s__elev%d__mp%d" % (name, elev, mp)
        wa, wv = g[key + "__analysis"].ravel(), g[key + "__variance"].ravel()
        sa, sv = max(2.0, float(np.abs(wa).max())), max(1.0, float(np.abs(wv).max()))
        assert_close(out, wa, sa, 1e-5, "oracle spatial " + key, allow_outliers=3)
        assert_close(out, wa, sa, 1e-3, "oracle spatial (loose) " + key)
        assert_close(var, wv, sv, 1e-5, "oracle spatial variance " + key, allow_outliers=3)
        assert_close(var, wv, sv, 1e-3, "oracle spatial variance (loose) " + key)


def test_spatially_varying_structures_live_against_reference(orc, ref):
    """Fresh inputs, horizontal scales only (well-conditioned systems): the restatement and the compiled reference agree
    to 1e-5 everywhere, for every family, limited and unlimited max_points, with and without extrapolation."""
    rng = np.random.default_rng(31)
    f32 = np.float32
    ny, nx, dx, S = 14, 17, 2000.0, 60
    yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    bg = rng.normal(size=(ny, nx)).astype(f32)
    pbg = rng.normal(size=S).astype(f32)
    obs = (pbg + rng.normal(size=S)).astype(f32)
    ratios = rng.uniform(0.3, 1.0, S).astype(f32)
    gy, gx = np.meshgrid(np.linspace(0, ny * dx, 4), np.linspace(0, nx * dx, 5), indexing="ij")
    h = rng.uniform(4000, 9000, size=(4, 5)).astype(f32)
    zero = np.zeros((4, 5), f32)
    bpts, opts = (yy, xx, None, None), (py, px, None, None)
    for stype in (B.BARNES, B.SOAR, B.TOAR, B.POWERLAW, B.LINEAR):
        for mp, extr, min_rho in ((8, True, 0.0013), (0, False, 0.1)):
            args = (bpts, bg, opts, obs, ratios, pbg, stype, (gy, gx), h, zero, zero, min_rho, mp, B.CARTESIAN)
            a, av = ref.optimal_interpolation_spatial(*args, allow_extrapolation=extr, want_variance=True)
            b, bv = orc.optimal_interpolation_spatial(*args, allow_extrapolation=extr, want_variance=True)
            assert_close(b, a, 1.0, 1e-5, "spatial family %d mp=%d" % (stype, mp))
            assert_close(bv, av, 1.0, 1e-5, "spatial variance family %d mp=%d" % (stype, mp))


STAT_CODES = dict(mean=B.MEAN, min=B.MIN, median=B.MEDIAN, max=B.MAX, std=B.STD, variance=B.VARIANCE, sum=B.SUM, count=B.COUNT)


def test_golden_statistics(orc):
    """Round 2: Std / Variance / Median neighbourhoods, the brute-force and exact-quantile neighbourhoods (vec2 and vec3),
    calc_statistic, calc_quantile and interpolate against the fixture generated from the compiled reference."""
    g = golden("statistics")
    f, e, rows = g["field"], g["ensemble"], g["rows"]
    for key in g.files:
        if key.startswith("nbh_hw"):
            hw, name = int(key[6:].split("__")[0]), key.split("__")[1]
            assert_bit_exact(orc.neighbourhood(f, hw, STAT_CODES[name]), g[key], key)
        elif key.startswith("brute_ens_hw") or key.startswith("brute_hw"):
            ens = key.startswith("brute_ens_hw")
            hw, name = int(key.split("hw")[1].split("__")[0]), key.split("__")[1]
            assert_bit_exact(orc.neighbourhood_window(e if ens else f, hw, STAT_CODES[name]), g[key], key)
        elif key.startswith("quantile_"):
            ens = key.startswith("quantile_ens_hw")
            hw, q = int(key.split("hw")[1].split("__")[0]), float(key.split("__q")[1])
            assert_bit_exact(orc.neighbourhood_window(e if ens else f, hw, B.QUANTILE, q), g[key], key)
        elif key.startswith("rows__q"):
            q = float(key[7:])
            assert_bit_exact(np.array([orc.calc_quantile(r, q) for r in rows], np.float32), g[key], key)
        elif key.startswith("rows__"):
            assert_bit_exact(np.array([orc.calc_statistic(r, STAT_CODES[key[6:]]) for r in rows], np.float32), g[key], key)
    got = np.array([orc.interpolate(v, g["interp_ix"], g["interp_iy"]) for v in g["interp_x"]], np.float32)
    assert_bit_exact(got, g["interp_y"], "interpolate")


def _gridding_cases(g):
    """(key, callable(lib) -> result) for every array of the gridding fixture; shared by the oracle and the GPU tests."""
    for tag, ctype in (("cart", B.CARTESIAN), ("geo", B.GEODETIC)):
        a = {k[len(tag) + 2:]: g[k] for k in g.files if k.startswith(tag + "__") and "__" not in k[len(tag) + 2:]}
        sets = dict(grid=(a["glats"], a["glons"]), points=(a["plats"], a["plons"]),
                    ogrid=(a["glats"][:9, :11] + np.float32(0.004 if ctype == B.GEODETIC else 400.0), a["glons"][:9, :11]),
                    opoints=(a["olats"], a["olons"]))
        for key in g.files:
            if not key.startswith(tag + "__") or key.count("__") < 2:
                continue
            _, func, spec = key.split("__")
            yield key, tag, ctype, a, sets, func, spec


def test_golden_gridding(orc):
    """Round 2: gridding / gridding_nearest / count / distance / fill / fill_missing / doping against the fixture
    generated from the compiled reference."""
    g = golden("gridding")
    n = 0
    for key, tag, ctype, a, sets, func, spec in _gridding_cases(g):
        radius = float(a["radius"])
        if func.startswith("gridding"):
            nearest = "nearest" in func
            oset = sets["grid"] if func.endswith("grid") else sets["opoints"]
            name = spec.split("_mn")[0]
            mn = int(spec.split("_mn")[1]) if "_mn" in spec else (0 if nearest else 1)
            got = orc.gridding(oset, sets["points"], a["values"], radius, mn, STAT_CODES[name], ctype, nearest=nearest)
        elif func == "count":
            i, o = spec.split("_")
            got = orc.count(sets[i], sets[o], radius, ctype)
        elif func == "distance":
            i, o, num = spec.split("_")
            got = orc.distance(sets[i], sets[o], int(num[1:]), ctype)
        elif func == "fill":
            got = orc.fill(a["glats"], a["glons"], a["field"], a["plats"], a["plons"], a["radii"], -7.5, int(spec[-1]), ctype)
        else:
            med = np.nan if spec == "nocheck" else 150.0
            square = func == "doping_square"
            got = orc.doping(a["glats"], a["glons"], a["gelevs"], a["field"], a["plats"], a["plons"], a["pelevs"], a["values"],
                             a["halfwidth"] if square else a["radii"], med, ctype, square)
        assert_bit_exact(got, g[key], key)
        n += 1
    assert n >= 100
    assert_bit_exact(orc.fill_missing(g["fill_missing__in"]), g["fill_missing__out"], "fill_missing")


def _ensi_multi_cases(g):
    """(tag, ctype, structure, arrays, [(key, kind, max_points, extrapolation, background, pbackground)]) of ensi_multi.npz"""
    for tag, ctype in (("cart", B.CARTESIAN), ("geo", B.GEODETIC)):
        spec = g[tag + "__structure"]
        a = {k[len(tag) + 2:]: g[k] for k in g.files if k.startswith(tag + "__")}
        runs = []
        for mp in (0, 12):
            for extr in (0, 1):
                for kind in ("ebesc", "ebe", "utem"):
                    runs.append(("%s_mp%d_x%d" % (kind, mp, extr), kind, mp, extr, a["background"], a["pbackground"]))
        for kind in ("ebesc", "ebe", "utem"):
            runs.append(("tail_" + kind, kind, 12, 0, a["tail_background"], a["tail_pbackground"]))
        yield tag, ctype, (int(spec[0]), float(spec[1]), float(spec[2]), float(spec[3])), a, runs


def test_golden_ensi_multi(orc):
    """Round 2: optimal_interpolation_ensi_multi_{ebe,ebesc,utem} (oi_ensi_multi.cpp:329-1311) and staticcorr_points
    (corr_points.cpp:26-131): the C restatement against the fixture generated from the compiled reference
    (tests/golden/make_golden_ensi_multi.py). The restatement follows the reference's operation order, so: bit for bit."""
    g = golden("ensi_multi")
    for tag, ctype, spec, a, runs in _ensi_multi_cases(g):
        s = B.make_structure(*spec)
        bp, op = (a["by"], a["bx"], a["be"], a["bf"]), (a["py"], a["px"], a["pe"], a["pf"])
        for mp in (0, 12):
            assert_bit_exact(orc.staticcorr_points(bp, op, s, mp, ctype), a["staticcorr_mp%d" % mp], "%s staticcorr mp=%d" % (tag, mp))
        for key, kind, mp, extr, bg, pbg in runs:
            corr = kind != "ebesc"
            got = orc.ensi_multi(kind, bp, a["bratios"], bg, a["background_corr"] if corr else None, op, a["pobs1"] if kind == "utem" else a["pobs2"],
                                 a["pratios"], pbg, a["pbackground_corr"] if corr else None, s, mp, ctype, bool(extr))
            assert_bit_exact(got, a[key], tag + " " + key)
    # an invalid member that is not the last one: the reference indexes past its innovation matrix (oi_ensi_multi.cpp:557,:806)
    tag, ctype, spec, a, runs = next(_ensi_multi_cases(g))
    bg = a["background"].copy()
    bg[0, 2] = np.nan
    with pytest.raises(RuntimeError):
        orc.ensi_multi("ebesc", (a["by"], a["bx"], a["be"], a["bf"]), a["bratios"], bg, None, (a["py"], a["px"], a["pe"], a["pf"]), a["pobs2"],
                       a["pratios"], a["pbackground"], None, B.make_structure(*spec), 12, ctype, True)


def test_golden_window_filters(orc):
    """Round 2: neighbourhood_search (neighbourhood_search.cpp:7-113) and calc_gradient (calc_gradient.cpp:6-126) restated
    in C against the fixture generated from the compiled reference (tests/golden/make_golden_window_filters.py): bit for bit."""
    g = golden("window_filters")
    for i, (hw, lo, hi, delta) in enumerate(g["search__cases"]):
        hw = int(hw)
        assert_bit_exact(orc.neighbourhood_search(g["search__array"], g["search__search"], hw, lo, hi, delta), g["search__case%d" % i], "search %d" % i)
        assert_bit_exact(orc.neighbourhood_search(g["search__array"], g["search__search"], hw, lo, hi, delta, g["search__apply"]),
                         g["search__case%d_apply" % i], "search %d with apply_array" % i)
    for i, (hw, num_min, min_range, default) in enumerate(g["gradient__cases"]):
        hw, num_min = int(hw), int(num_min)
        for name, gt in (("minmax", 0), ("regression", 10)):
            assert_bit_exact(orc.calc_gradient(g["gradient__base"], g["gradient__values"], gt, hw, num_min, min_range, default),
                             g["gradient__%s_case%d" % (name, i)], "gradient %s %d" % (name, i))
    with pytest.raises(ValueError):
        orc.neighbourhood_search(g["search__array"], g["search__search"], 1, 2.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        orc.calc_gradient(g["gradient__base"], g["gradient__values"], 0, 0)


def test_ensi_oracle_against_lapack(orc):
    """The EnSI numerics of both checkers rest on hand-written dense routines (Gauss-Jordan `inv`, exact 1-norm `rcond`, cyclic
    Jacobi `eig_sym`: oracle/shims/armadillo and oracle/gridpp_oracle.c). This restates oi_ensi.cpp:282-554 per grid point with
    LAPACK underneath (numpy.linalg.inv -> dgetrf/dgetri, numpy.linalg.eigh -> dsyevd) on the golden case and compares with the
    oracle: the selection and rho come from the oracle's own structure function, the dense algebra does not. 1e-6 relative."""
    g = golden("ensi_c5_density")
    s = oracle_structure(parse_spec(g["structure"]))
    y, x, bg = g["y"].ravel(), g["x"].ravel(), g["background"].reshape(-1, g["background"].shape[-1])
    py, px, pobs, psig, pbg = g["py"], g["px"], g["pobs"], g["psigmas"], g["pbackground"]
    nB, E, S = bg.shape[0], bg.shape[1], py.size
    R = orc.structure_localization_distance(s)
    nanv = np.full(S, np.nan, f32)
    obs_pts = np.stack([px, py, np.zeros(S, f32), nanv, nanv], axis=1)
    # :163-178 (all members valid in the fixture); float accumulation in member order, as calc_statistic does
    gYhat = np.array([np.float32(sum((np.float32(v) for v in row), np.float32(0))) / np.float32(E) for row in pbg], f32)
    gY = (pbg - gYhat[:, None]).astype(f32)
    for name in ("mp20", "mp20_clamp", "unlimited"):
        mp, extr = (int(v) for v in g[name + "__args"])
        want = g[name + "__analysis"].reshape(nB, E)
        got = bg.copy()
        for b in range(nB):
            d = np.sqrt((px - x[b]) * (px - x[b]) + (py - y[b]) * (py - y[b]), dtype=f32)
            near = np.nonzero(d <= R)[0]
            if near.size == 0:
                continue
            p1 = np.tile(np.array([x[b], y[b], 0, np.nan, np.nan], f32), (near.size, 1))
            rho = orc.structure_corr(s, p1, obs_pts[near], background=True)
            keep = (rho > 0) & np.isfinite(pobs[near])                          # :232-236: only pobs validity is tested
            idx, rho = near[keep], rho[keep]
            if mp > 0 and idx.size > mp:                                       # :244-255: best first
                order = np.lexsort((idx, -rho.astype(np.float64)))[:mp]
                idx, rho = idx[order], rho[order]
            if idx.size == 0:
                continue
            lY = gY[idx].astype(np.float64)                                    # lS x E
            rinv = rho.astype(np.float64) / (psig[idx] * psig[idx]).astype(np.float64)
            Cm = lY.T * rinv[None, :]
            Pinv = Cm @ lY + float(f32(E - 1)) * np.eye(E)
            if 1.0 / np.linalg.cond(Pinv, 1) <= 0:
                continue
            P = np.linalg.inv(Pinv)
            val, vec = np.linalg.eigh((E - 1) * P)
            W = (vec * np.sqrt(val)[None, :]) @ vec.T
            w = P @ Cm @ (pobs[idx].astype(np.float64) - gYhat[idx].astype(np.float64))
            W = W + w[:, None]
            total = np.float32(0)
            for v in bg[b]:
                total = np.float32(total + v)
            ens_mean = np.float32(total / np.float32(E))
            X = bg[b].astype(np.float64) - float(ens_mean)
            for e in range(E):
                tot = np.float32(0)
                for k in range(E):
                    tot = np.float32(float(tot) + X[k] * W[k, e])
                inc = tot
                if not extr:
                    lYe = lY.ravel(order="F")[e]                                   # the linear index lY[e] of :523-524
                    dd = pobs[idx].astype(np.float64) - (lYe + gYhat[idx].astype(np.float64))
                    max_inc, min_inc = np.float32(dd.max()), np.float32(dd.min())
                    member = np.float32(float(inc) - X[e])
                    if max_inc > 0 and member > max_inc:
                        inc = np.float32(float(max_inc) + X[e])
                    elif max_inc < 0 and member > 0:
                        inc = np.float32(X[e])
                    elif min_inc < 0 and member < min_inc:
                        inc = np.float32(float(min_inc) + X[e])
                    elif min_inc > 0 and member < 0:
                        inc = np.float32(X[e])
                got[b, e] = np.float32(ens_mean + inc)
        assert np.abs(want - bg).max() > 0.05
        assert_close(got, want, 1.0, 1e-6, "LAPACK restatement, " + name)


def test_ensi_multi_oracle_against_lapack(orc):
    """ebesc / ebe (oi_ensi_multi.cpp:630-859, :329-627) per point on numpy.linalg (LAPACK `inv`), to take the hand-written
    Gauss-Jordan of both checkers out of the loop: selection and correlations from the oracle's own structure function
    (staticcorr_points on the observations with a valid value), the dense algebra from LAPACK. 1e-6 relative."""
    g = golden("ensi_multi")
    tag, ctype, spec, a, runs = next(_ensi_multi_cases(g))
    s = B.make_structure(*spec)
    E = a["background"].shape[1]
    ok = np.isfinite(a["pobs2"][:, 0])
    vi = np.nonzero(ok)[0]
    bp = (a["by"], a["bx"], a["be"], a["bf"])
    op = tuple(a[k][vi] for k in ("py", "px", "pe", "pf"))
    nV = vi.size
    # observation-observation correlations, all pairs at once: structure.corr(p_i, p_j) on (x, y, z, elev, laf) rows (Cartesian)
    P = np.stack([op[1], op[0], np.zeros(nV, f32), op[2], op[3]], axis=1)
    Call = orc.structure_corr(s, np.repeat(P, nV, axis=0), np.tile(P, (nV, 1))).reshape(nV, nV).astype(np.float64)
    innov = (a["pobs2"][vi] - a["pbackground"][vi]).astype(f32).astype(np.float64)

    def standardise(rows):
        out = np.zeros(rows.shape, np.float64)
        for i, row in enumerate(rows):
            mean = np.float32(sum((np.float32(v) for v in row), np.float32(0)) / np.float32(E))
            d = (row - row[0]).astype(f32)
            m1 = np.float32(sum((np.float32(v) for v in d), np.float32(0)) / np.float32(E))
            m2 = np.float32(sum((np.float32(v * v) for v in d), np.float32(0)) / np.float32(E))
            sd = np.float32(np.sqrt(max(np.float32(m2 - m1 * m1), np.float32(0))))
            if sd > np.float32(0.0013):
                out[i] = 1 / np.sqrt(E - 1.0) * (row - mean).astype(np.float64) / float(sd)
        return out
    Z = standardise(a["pbackground_corr"][vi]).astype(f32).astype(np.float64)       # gZ_R is a float table in the reference
    XL = standardise(a["background_corr"])
    for mp, extr in ((12, 0), (0, 1)):
        rho = orc.staticcorr_points(bp, op, s, mp, ctype).astype(np.float64)       # (L, nV): the selection and its correlations
        for kind in ("ebesc", "ebe"):
            got = a["background"].copy()
            for y in range(rho.shape[0]):
                sel = np.nonzero(rho[y] > 0)[0]
                if sel.size == 0:
                    continue
                A = Call[np.ix_(sel, sel)].copy()
                r = rho[y, sel].copy()
                if kind == "ebe":
                    A *= Z[sel] @ Z[sel].T
                    r *= Z[sel] @ XL[y]
                A[np.diag_indices(sel.size)] += a["pratios"][vi][sel].astype(np.float64)
                K = r @ np.linalg.inv(A)
                dx = float(a["bratios"][y]) * (K @ innov[sel])
                if not extr:
                    mx, mn = innov[sel].max(axis=0).astype(f32), innov[sel].min(axis=0).astype(f32)
                    # the reference's if / else-if chain: the first condition that holds wins
                    first = (mx > 0) & (dx.astype(f32) > mx)
                    second = ~first & (mx < 0) & (dx.astype(f32) > 0)
                    third = ~first & ~second & (mn < 0) & (dx.astype(f32) < mn)
                    fourth = ~first & ~second & ~third & (mn > 0) & (dx.astype(f32) < 0)
                    inc = np.where(first, mx, np.where(second | fourth, np.float32(0), np.where(third, mn, dx.astype(f32)))).astype(f32)
                    dx = inc.astype(np.float64)
                got[y] = (a["background"][y].astype(np.float64) + dx).astype(f32)
            want = a["%s_mp%d_x%d" % (kind, mp, extr)]
            assert_close(got, want, 1.0, 1e-6, "LAPACK restatement of %s mp=%d extrapolation=%d" % (kind, mp, extr), allow_outliers=2 if not extr else 0)
