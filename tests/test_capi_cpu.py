"""CPU-side checks of the product library: it loads, exports every symbol include/gridpp_b200.h declares, its
host-only parts (structure descriptors, coordinate conversion, argument validation) agree with the oracle, and
every compute entry point FAILS LOUDLY without a GPU (there is no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from oracle import bindings as B
from util import assert_bit_exact, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def _has_gpu(gpp):
    return gpp.device_count() > 0


def test_header_symbols_are_exported(gpp):
    header = open(os.path.join(ROOT, "include", "gridpp_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(gpp_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    from gridpp_b200 import _lib
    for name in declared:
        assert hasattr(_lib.lib, name), "libgridpp_b200.so does not export " + name
    assert sorted(_lib.EXPORTS) == declared, "python prototypes and header disagree: %s" % (set(_lib.EXPORTS) ^ set(declared))


def test_no_oracle_or_cpu_path_in_product():
    """The product must not import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gridpp_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text and "libgridpp_ref" not in text, fn
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), fn


def test_structure_descriptors_match_oracle(gpp, orc):
    classes = {B.BARNES: gpp.BarnesStructure, B.CRESSMAN: gpp.CressmanStructure, B.SOAR: gpp.SoarStructure,
               B.TOAR: gpp.ToarStructure, B.POWERLAW: gpp.PowerlawStructure, B.LINEAR: gpp.LinearStructure}
    rng = np.random.default_rng(3)
    for st, cls in classes.items():
        for h in list(rng.uniform(100, 1e5, 40).astype(f32)) + [2000.0, 10000.0]:
            for hmax in (float("nan"), 0.0, 0.7 * float(h), 3.0 * float(h)):
                s = cls(float(h), 100.0, 0.5) if st == B.CRESSMAN else cls(float(h), 100.0, 0.5, hmax)
                want = orc.structure_describe(st, float(h), 100.0, 0.5, hmax)
                got = s.localization_distance()
                assert f32(got) == f32(want) or (np.isnan(got) and np.isnan(want)), (st, h, hmax, got, want)
                o = B.make_structure(st, float(h), 100.0, 0.5, hmax)
                assert f32(s._desc.term[0].min_rho) == f32(o.term[0].min_rho), (st, h, hmax)
    # tests/test_structure.py:8-62 of the reference: constructor validation -> ValueError
    for cls in (gpp.BarnesStructure, gpp.CressmanStructure):
        for bad in (-1, np.nan):
            with pytest.raises(ValueError):
                cls(bad)
            with pytest.raises(ValueError):
                cls(2000, bad)
            with pytest.raises(ValueError):
                cls(2000, 100, bad)
    with pytest.raises(ValueError):
        gpp.BarnesStructure(2000, 100, 0, -1)
    for dist in (-1, np.nan):
        with pytest.raises(ValueError):
            gpp.CrossValidation(gpp.BarnesStructure(2000), dist)
    g = golden("structure")
    assert f32(gpp.BarnesStructure(5000, 200, 0.5).localization_distance()) == g["barnes__loc_dist"]
    m = gpp.MultipleStructure(gpp.BarnesStructure(5000), gpp.CressmanStructure(300, 300, 300), gpp.LinearStructure(0.3, 0.3, 0.3))
    assert f32(m.localization_distance()) == g["multiple__loc_dist"]
    assert f32(m.clone().localization_distance()) == g["multiple__loc_dist"]


def test_points_coordinates_match_reference(gpp):
    g = golden("index_queries")
    for tname, t in (("geodetic", gpp.Geodetic), ("cartesian", gpp.Cartesian)):
        p = gpp.Points(g[tname + "__lats"], g[tname + "__lons"], type=t)
        for got, want in zip(p._set.xyz(), (g[tname + "__x"], g[tname + "__y"], g[tname + "__z"])):
            assert_bit_exact(got, want, tname)
        assert p.size() == 1500 and p.get_coordinate_type() == t
        assert np.isnan(p.get_elevs()).all() and np.isnan(p.get_lafs()).all()     # points.cpp:23-30
    # tests/test_kdtree.py:167-186 of the reference
    for lat, lon in ((91, 0), (-91, 0), (np.nan, 0), (0, np.nan)):
        with pytest.raises(ValueError):
            gpp.KDTree([lat], [lon], gpp.Geodetic)
    gpp.KDTree([0, 90.000001], [0, 0])
    with pytest.raises(ValueError):
        gpp.Points([0, 1], [0])
    with pytest.raises(ValueError):
        gpp.Points([0, 1], [0, 1], [1])
    grid = gpp.Grid([[0, 0], [1, 1], [2, 2]], [[0, 1], [0, 1], [0, 1]])
    assert grid.size().tolist() == [3, 2] and grid.to_points().size() == 6
    assert gpp.Grid().size().tolist() == [0, 0] and gpp.Points().size() == 0


def test_argument_validation(gpp):
    """tests/test_optimal_interpolation.py:9-47 of the reference: every size mismatch raises ValueError (checked
    before any device work, so this runs without a GPU)."""
    y, x = np.meshgrid(np.arange(3) * 1000.0, np.arange(4) * 1000.0, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    points = gpp.Points([0, 1000], [0, 1000], type=gpp.Cartesian)
    s = gpp.BarnesStructure(2500)
    ok = dict(bgrid=grid, background=np.zeros((3, 4)), points=points, pobs=[1, 2], pratios=[0.5, 0.5], pbackground=[0, 0],
              structure=s, max_points=5)
    bad = dict(background=[np.zeros((3, 3)), np.zeros((2, 4)), np.zeros(12)], pobs=[[1], [1, 2, 3]], pratios=[[1], [1, 2, 3]],
               pbackground=[[1], [1, 2, 3]], max_points=[-1], points=[gpp.Points([0, 1], [0, 1])])
    for key, values in bad.items():
        for v in values:
            args = dict(ok)
            args[key] = v
            with pytest.raises(ValueError):
                gpp.optimal_interpolation(*args.values())
    with pytest.raises(ValueError):
        gpp.neighbourhood(np.ones((5, 5)), -1, gpp.Mean)
    with pytest.raises(ValueError):
        gpp.neighbourhood(np.ones((5, 5)), 1, gpp.Quantile)
    with pytest.raises(ValueError):
        gpp.neighbourhood_quantile_fast(np.ones((5, 5)), 0.5, -1, [0, 1])
    assert gpp.neighbourhood([[]], 1, gpp.Mean).shape == (0, 0)
    assert gpp.neighbourhood_quantile_fast([[]], 0.9, 1, [0, 1]).shape == (0, 0)
    with pytest.raises(ValueError):
        gpp.nearest(grid, points, np.zeros((2, 2)))


def test_compute_fails_loudly_without_gpu(gpp):
    if _has_gpu(gpp):
        pytest.skip("a GPU is present; the no-fallback behaviour is only observable on a CPU box")
    y, x = np.meshgrid(np.arange(3) * 1000.0, np.arange(4) * 1000.0, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    points = gpp.Points([0, 1000], [0, 1000], type=gpp.Cartesian)
    calls = [
        lambda: gpp.neighbourhood(np.ones((4, 4)), 1, gpp.Mean),
        lambda: gpp.neighbourhood_quantile_fast(np.ones((4, 4)), 0.5, 1, [0, 1]),
        lambda: gpp.optimal_interpolation(grid, np.zeros((3, 4)), points, [1, 2], [0.5, 0.5], [0, 0], gpp.BarnesStructure(2500), 5),
        lambda: gpp.nearest(grid, points, np.zeros((3, 4))),
        lambda: points.get_nearest_neighbour(0, 0),
        lambda: gpp.BarnesStructure(2500).corr([[0, 0, 0, 0, 0]], [[0, 1, 0, 0, 0]]),
    ]
    for call in calls:
        with pytest.raises(RuntimeError, match="no usable CUDA device"):
            call()


def test_small_integer_quotients_are_exact():
    """csrc/quantile_tma.cu evaluates F = fl(n / m) (0 <= n <= m <= 31^2) as q0 = n * r, q = fma(fma(-q0, m, n), r, q0)
    with r = fl(1 / m) instead of a float division. Exhaustive check that this is the correctly rounded quotient (the
    value the reference forms, neighbourhood.cpp:374-390). The float64 emulation of the two fma's is exact here: the
    products have <= 48 significant bits, and n / m with m < 2^10 is never within 2^-44 of a rounding boundary."""
    f64 = np.float64
    for m in range(1, 31 * 31 + 1):
        n = np.arange(0, m + 1)
        fm = f32(m)
        r = f32(1) / fm
        fn = n.astype(f32)
        q0 = fn * r
        e = (fn.astype(f64) - q0.astype(f64) * f64(fm)).astype(f32)
        q = (e.astype(f64) * f64(r) + q0.astype(f64)).astype(f32)
        want = (n.astype(f64) / f64(m)).astype(f32)
        assert (q == want).all(), m


def test_convert_coordinates_and_min_rho_constructors(gpp, orc):
    """The two host-only entry points the C++ layer needs: gpp_convert_coordinates (util.cpp:583-615, what Point's
    constructor calls) against the oracle bit for bit, and gpp_structure_init_min_rho (the (1, 1)-field form of
    <Family>Structure(Grid, h, v, w, min_rho), structure.cpp:168-176) against localization_distance(h) of each family."""
    import ctypes as C
    from gridpp_b200 import _lib
    rng = np.random.default_rng(17)
    lats, lons = rng.uniform(-90, 90, 500).astype(f32), rng.uniform(-360, 360, 500).astype(f32)
    for ctype in (gpp.Geodetic, gpp.Cartesian):
        got = [np.empty(500, f32) for _ in range(3)]
        rc = _lib.lib.gpp_convert_coordinates(lats.ctypes.data_as(_lib.fp), lons.ctypes.data_as(_lib.fp), 500, ctype,
                                              *[a.ctypes.data_as(_lib.fp) for a in got])
        assert rc == 0
        for a, b, axis in zip(got, orc.convert_coordinates(lats, lons, ctype), "xyz"):
            assert_bit_exact(a, b, "convert_coordinates " + axis)
    bad = np.array([95.0], f32)
    out = [np.empty(1, f32) for _ in range(3)]
    rc = _lib.lib.gpp_convert_coordinates(bad.ctypes.data_as(_lib.fp), bad.ctypes.data_as(_lib.fp), 1, gpp.Geodetic,
                                          *[a.ctypes.data_as(_lib.fp) for a in out])
    assert rc == 1 and b"Invalid coords" in _lib.lib.gpp_last_error()      # util.cpp:596-600 -> std::invalid_argument
    # tests/test_barnes_structure.py:46-66 of the reference: sqrt(-2 log(0.1)) * 2500
    grid = gpp.Grid([[0.0]], [[0.0]], type=gpp.Cartesian)
    s = gpp.BarnesStructure(grid, [[2500]], [[0]], [[0]], 0.1)
    assert abs(s.localization_distance() - np.sqrt(-2 * np.log(0.1)) * 2500) < 1e-2
    for st, cls in ((B.BARNES, gpp.BarnesStructure), (B.SOAR, gpp.SoarStructure), (B.TOAR, gpp.ToarStructure),
                    (B.POWERLAW, gpp.PowerlawStructure), (B.LINEAR, gpp.LinearStructure)):
        for h, min_rho in ((2500.0, 0.1), (10000.0, 0.0013), (731.5, 0.4)):
            got = cls(grid, [[h]], [[100]], [[0.5]], min_rho).localization_distance()
            o = B.make_structure(st, h, 100.0, 0.5, min_rho=min_rho)
            want = C.c_float()
            orc._check(orc._fn("structure_localization_distance")(C.byref(o), C.byref(want)))
            assert f32(got) == f32(want.value), (st, h, min_rho, got, want.value)
    d = _lib.StructureDesc()
    assert _lib.lib.gpp_structure_init_min_rho(C.byref(d), B.CRESSMAN, 1000.0, 0.0, 0.0, 0.1) == 1   # no such constructor


def test_point_constructor_follows_the_reference(gpp, orc):
    """gridpp::Point (point.cpp:5-26): geodetic points carry the converted coordinates (util.cpp:583-615, invalid latitudes
    throw), Cartesian ones x = lat, y = lon without any check; the eight-argument form stores what it is given."""
    p = gpp.Point(60, 10, 100, 0.5)
    x, y, z = orc.convert_coordinates([60], [10], gpp.Geodetic)
    assert (f32(p.x), f32(p.y), f32(p.z)) == (x[0], y[0], z[0]) and p.elev == 100 and p.laf == 0.5 and p.type == gpp.Geodetic
    q = gpp.Point(1, 2, 0, 0, gpp.Cartesian)
    assert (q.x, q.y, q.z) == (1, 2, 0)
    assert np.isnan(gpp.Point(np.nan, 0, 0, 0, gpp.Cartesian).x) and np.isnan(gpp.Point(1, 2).elev)
    with pytest.raises(ValueError):
        gpp.Point(95, 0)
    r = gpp.Point(1, 2, 3, 4, gpp.Cartesian, 7, 8, 9)
    assert (r.x, r.y, r.z, r.lat, r.lon) == (7, 8, 9, 1, 2)
    # Point arguments become rows of (x, y, z, elev, laf); a spatially varying structure function needs a Point for p1
    rows = gpp.StructureFunction._points5([q, r])
    assert rows.shape == (2, 5) and rows[1].tolist() == [7, 8, 9, 3, 4]
    grid = gpp.Grid([[0, 0]], [[0, 2500]], type=gpp.Cartesian)
    s = gpp.BarnesStructure(grid, [[2500, 1]], [[0, 0]], [[0, 0]], 0.1)
    with pytest.raises(ValueError):
        s.corr(np.zeros((1, 5)), np.zeros((1, 5)))


def test_scalar_helpers_match_reference(gpp, orc):
    """KDTree::calc_distance / calc_straight_distance / deg2rad / rad2deg (kdtree.cpp:107-200), also under their flat SWIG
    names, is_valid (util.cpp:16-18) and the thread-count functions (gridpp.cpp:45-68)."""
    rng = np.random.default_rng(23)
    for ctype, lo, hi in ((gpp.Geodetic, -90, 90), (gpp.Cartesian, -1e5, 1e5)):
        for _ in range(300):
            a, c = rng.uniform(lo, hi, 2).astype(f32)
            b, d = rng.uniform(2 * lo, 2 * hi, 2).astype(f32)
            want = orc.calc_distance(float(a), float(b), float(c), float(d), ctype)
            got = gpp.KDTree.calc_distance(a, b, c, d, ctype)
            assert f32(got) == f32(want), (ctype, a, b, c, d, got, want)
    # tests/test_kdtree.py:108-112 and :50-58 of the reference
    assert abs(gpp.KDTree_calc_distance(0, 0, 0.001, 0.001) - 157.42953491210938) < 1e-5
    p0, p1 = gpp.Point(0, 0), gpp.Point(0.001, 0.001)
    assert abs(gpp.KDTree_calc_straight_distance(p0.x, p0.y, p0.z, p1.x, p1.y, p1.z) - 157.42953491210938) < 0.5
    assert gpp.KDTree.calc_distance(p0, p1) == gpp.KDTree_calc_distance(0, 0, 0.001, 0.001)
    assert gpp.KDTree.calc_distance(60, 10, 60, 10) == 0
    assert abs(gpp.KDTree_rad2deg(1) - 180 / 3.14159265) < 1e-5 and abs(gpp.KDTree_deg2rad(180) - 3.14159265) < 1e-5
    with pytest.raises(RuntimeError):
        gpp.KDTree.calc_distance(gpp.Point(0, 0), gpp.Point(0, 0, 0, 0, gpp.Cartesian))
    assert gpp.is_valid(1.0) and not gpp.is_valid(float("nan")) and not gpp.is_valid(float("inf"))
    gpp.set_omp_threads(3)
    assert gpp.get_omp_threads() == 3
    gpp.initialize_omp()
    tree = gpp.KDTree([60, 61], [10, 11])
    x, y, z = orc.convert_coordinates([60, 61], [10, 11], gpp.Geodetic)
    assert_bit_exact(tree.get_x(), x, "get_x")
    assert_bit_exact(tree.get_z(), z, "get_z")


def test_get_point_and_convert_coordinates(gpp):
    """tests/test_points.py:115-134 and tests/test_grid.py:80-90 of the reference."""
    points = gpp.Points([0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11])
    p = points.get_point(1)
    assert (p.lat, p.lon, p.elev, p.laf) == (1, 4, 7, 10)
    s, x, y, z = gpp.convert_coordinates(1, 4, gpp.Geodetic)
    assert s and (p.x, p.y, p.z) == (x, y, z)
    grid = gpp.Grid([[0, 0], [1, 1]], [[0, 0.25], [0, 0.25]], [[1, 2], [3, 4]], [[0, 0.1], [0.2, 0.3]])
    q = grid.get_point(1, 1)
    s, x, y, z = gpp.convert_coordinates(1, 0.25, gpp.Geodetic)
    assert (q.lat, q.lon, q.elev) == (1, 0.25, 4) and abs(q.laf - 0.3) < 1e-7 and (q.x, q.y, q.z) == (x, y, z)
    s, xs, ys, zs = gpp.convert_coordinates([0, 1], [3, 4], gpp.Cartesian)
    assert xs.tolist() == [3, 4] and ys.tolist() == [0, 1] and zs.tolist() == [0, 0]
    with pytest.raises(ValueError):
        points.get_point(3)
    with pytest.raises(ValueError):
        gpp.convert_coordinates(91, 0, gpp.Geodetic)


def test_spatial_structure_descriptor_is_refused_by_plain_entry_points(gpp):
    """the placeholder descriptor of a spatially varying <Family>Structure(grid, h, v, w) carries NaN
    scales, and every entry point that takes a plain descriptor refuses it (before touching the device) instead of
    analysing with a localization distance of zero."""
    from gridpp_b200._lib import NotImplementedOnDevice
    y, x = np.meshgrid(np.arange(6, dtype=f32) * 1000, np.arange(7, dtype=f32) * 1000, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    spatial = gpp.BarnesStructure(grid, np.full((6, 7), 2500, f32), np.zeros((6, 7), f32), np.zeros((6, 7), f32))
    assert np.isnan(spatial._desc.term[0].h) and np.isnan(spatial._desc.term[0].loc_dist)
    points = gpp.Points([1000, 3000], [1000, 4000], type=gpp.Cartesian)
    with pytest.raises(NotImplementedOnDevice):
        gpp.optimal_interpolation_ensi(grid, np.zeros((6, 7, 5), f32), points, [1.0, 2.0], [0.5, 0.5], np.zeros((2, 5), f32), spatial, 10)
    from gridpp_b200 import device as gd
    with pytest.raises(NotImplementedOnDevice):
        gd.ObservationState(points, [1.0, 2.0], [0.5, 0.5], [0.0, 0.0], spatial)
    with pytest.raises(NotImplementedOnDevice):
        gd.EnsembleObservationState(points, [1.0, 2.0], [0.5, 0.5], np.zeros((2, 5), f32), spatial)
