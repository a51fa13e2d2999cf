"""GPU parity tests added in round 2: the work-unit scheduler and launch workspace of the register OI kernel, the
device-resident EnSI entry point, and configs 4 and 5 of BASELINE.json at their full sizes."""
import numpy as np
import pytest

from oracle import bindings as B
from util import assert_bit_exact, assert_close

pytestmark = pytest.mark.gpu
f32 = np.float32
RTOL = 1e-5


def _c3_like(rng, ny, nx, dx, density=0.01):
    """C3 geometry on an ny x nx grid: `density` observations per km^2 over the area the grid can reach."""
    y, x = np.meshgrid(np.arange(ny, dtype=f32) * dx, np.arange(nx, dtype=f32) * dx, indexing="ij")
    pad = 40000.0
    S = int(density * (ny * dx + 2 * pad) * (nx * dx + 2 * pad) / 1e6)
    py = rng.uniform(-pad, ny * dx + pad, S).astype(f32)
    px = rng.uniform(-pad, nx * dx + pad, S).astype(f32)
    bg = (rng.standard_normal((ny, nx)) * 3).astype(f32)
    pbg = (rng.standard_normal(S) * 3).astype(f32)
    obs = (pbg + rng.standard_normal(S) * 0.5).astype(f32)
    return y, x, py, px, bg, pbg, obs, np.full(S, 0.5, f32)


def test_oi_work_units_workspace_and_traversal(gpp, orc):
    """The register kernel hands work out in units of 16, 4 and 1 tiles (the last rows of a range in the small ones) from
    a counter in a launch workspace that the kernel itself resets. A 700 x 704 grid has all three unit sizes. Checked:
    a sample against the oracle; the Grid traversal (4 x 4 tiles) against the Points traversal (16-point runs) and
    against ranges cut at arbitrary rows -- bit for bit; repeated launches on the same observation state (the workspace
    must come back clean); the explicit-workspace entry point."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(77)
    ny, nx, dx = 700, 704, 250.0
    y, x, py, px, bg, pbg, obs, ratios = _c3_like(rng, ny, nx, dx)
    bg[rng.uniform(size=bg.shape) < 0.001] = np.nan
    grid, points, s = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    state = gd.ObservationState(points, obs, ratios, pbg, s)
    d_bg = torch.from_numpy(bg.ravel()).cuda()
    whole = gd.optimal_interpolation(grid, d_bg, state, 30)
    pick = rng.choice(ny * nx, 5000, replace=False)
    pick[:4] = [0, nx - 1, nx * (ny - 1), ny * nx - 1]
    want = orc.optimal_interpolation((y.ravel()[pick], x.ravel()[pick], None, None), bg.ravel()[pick], (py, px, None, None), obs, ratios,
                                     pbg, B.make_structure(B.BARNES, 10000.0), 30, B.CARTESIAN)
    assert_close(whole.cpu().numpy()[pick], want, 3.0, RTOL, "work units vs oracle")
    # the same launch again and again: 7 launches cycle through the 4 workspace slots of the state
    for _ in range(7):
        again = gd.optimal_interpolation(grid, d_bg, state, 30)
    assert torch.equal(torch.nan_to_num(whole, nan=-7.0), torch.nan_to_num(again, nan=-7.0))
    # ranges cut inside tiles and units (a range that is not whole rows takes the 16-point-run traversal)
    parts = torch.full_like(d_bg, -1.0)
    cuts = [0, 3 * nx, 3 * nx + 5, 250 * nx, 251 * nx, 500 * nx + 17, ny * nx]
    for a, b in zip(cuts[:-1], cuts[1:]):
        gd.optimal_interpolation(grid, d_bg, state, 30, out=parts, first=a, count=b - a)
    assert torch.equal(torch.nan_to_num(whole, nan=-7.0), torch.nan_to_num(parts, nan=-7.0))
    # the Points traversal of the same nodes
    as_points = gpp.Points(y.ravel(), x.ravel(), type=gpp.Cartesian)
    runs = gd.optimal_interpolation(as_points, d_bg, state, 30)
    assert torch.equal(torch.nan_to_num(whole, nan=-7.0), torch.nan_to_num(runs, nan=-7.0))
    # caller-owned workspace
    ws = gd.Workspace()
    assert ws.bytes > 0
    for _ in range(2):
        mine = gd.optimal_interpolation_ws(grid, d_bg, state, 30, ws)
    assert torch.equal(torch.nan_to_num(whole, nan=-7.0), torch.nan_to_num(mine, nan=-7.0))
    # small and ragged shapes: every unit is a single tile, partial tiles on both edges
    for (sy, sx) in ((1, 1), (3, 5), (9, 130), (33, 37)):
        sub = (slice(0, sy), slice(0, sx))
        got = gpp.optimal_interpolation(gpp.Grid(y[sub], x[sub], type=gpp.Cartesian), bg[sub], points, obs, ratios, pbg, s, 30)
        assert_bit_exact(got, whole.cpu().numpy().reshape(ny, nx)[sub], "sub-grid %dx%d" % (sy, sx))


def test_oi_with_elevations_and_geodetic(gpp, orc):
    """The inlined Barnes evaluation with the vertical and land/sea terms active, on a Geodetic grid (three coordinate
    planes are read) and on a Cartesian one with elevations only."""
    rng = np.random.default_rng(5)
    ny, nx = 60, 70
    lats, lons = np.meshgrid(np.linspace(59, 60, ny).astype(f32), np.linspace(10, 12, nx).astype(f32), indexing="ij")
    elev = rng.uniform(0, 800, (ny, nx)).astype(f32)
    laf = rng.uniform(0, 1, (ny, nx)).astype(f32)
    elev[rng.uniform(size=elev.shape) < 0.02] = np.nan
    S = 400
    pla, plo = rng.uniform(58.9, 60.1, S).astype(f32), rng.uniform(9.8, 12.2, S).astype(f32)
    pel, plaf = rng.uniform(0, 800, S).astype(f32), rng.uniform(0, 1, S).astype(f32)
    pel[rng.uniform(size=S) < 0.05] = np.nan
    bg = rng.standard_normal((ny, nx)).astype(f32)
    pbg = rng.standard_normal(S).astype(f32)
    obs = (pbg + rng.standard_normal(S)).astype(f32)
    ratios = rng.uniform(0.2, 1.0, S).astype(f32)
    for mp, extr in ((30, True), (12, False)):
        got = gpp.optimal_interpolation(gpp.Grid(lats, lons, elev, laf), bg, gpp.Points(pla, plo, pel, plaf), obs, ratios, pbg,
                                        gpp.BarnesStructure(30000, 300, 0.5), mp, extr)
        want = orc.optimal_interpolation((lats, lons, elev, laf), bg, (pla, plo, pel, plaf), obs, ratios, pbg,
                                         B.make_structure(B.BARNES, 30000.0, 300.0, 0.5), mp, B.GEODETIC, allow_extrapolation=extr)
        assert_close(got.ravel(), want, 1.0, RTOL, "geodetic OI with elevation and laf terms mp=%d" % mp)
    y, x = np.meshgrid(np.arange(ny, dtype=f32) * 1000, np.arange(nx, dtype=f32) * 1000, indexing="ij")
    py, px = rng.uniform(0, ny * 1000, S).astype(f32), rng.uniform(0, nx * 1000, S).astype(f32)
    got = gpp.optimal_interpolation(gpp.Grid(y, x, elev, type=gpp.Cartesian), bg, gpp.Points(py, px, pel, type=gpp.Cartesian), obs, ratios,
                                    pbg, gpp.BarnesStructure(15000, 200), 30)
    want = orc.optimal_interpolation((y, x, elev, None), bg, (py, px, pel, None), obs, ratios, pbg,
                                     B.make_structure(B.BARNES, 15000.0, 200.0), 30, B.CARTESIAN)
    assert_close(got.ravel(), want, 1.0, RTOL, "cartesian OI with elevation term")


def test_ensi_device_entry_point(gpp, orc):
    """gpp_optimal_interpolation_ensi_device (observation state built once, device-resident ensemble, one launch per
    range) against the host entry point and the oracle; ranges, in-place analysis and the valid-member flags."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(21)
    ny, nx, dx, E, S = 48, 64, 200.0, 20, 900
    y, x = np.meshgrid(30000 + np.arange(ny, dtype=f32) * dx, 30000 + np.arange(nx, dtype=f32) * dx, indexing="ij")
    py, px = rng.uniform(0, 72000, S).astype(f32), rng.uniform(0, 76000, S).astype(f32)
    bg = (rng.normal(size=(ny, nx, 1)) * 2 + rng.normal(size=(ny, nx, E))).astype(f32)
    pbg = rng.normal(size=(S, E)).astype(f32)
    obs, sig = rng.normal(size=S).astype(f32), np.full(S, 0.5, f32)
    obs[3] = np.nan
    grid, points, s = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    host = gpp.optimal_interpolation_ensi(grid, bg, points, obs, sig, pbg, s, 50)
    d_bg = torch.from_numpy(bg).cuda()
    flags = gd.valid_members(d_bg)
    assert flags.all()
    state = gd.EnsembleObservationState(points, obs, sig, pbg, s, member_valid=flags)
    skipped = torch.zeros(1, dtype=torch.int32, device="cuda")
    whole = gd.optimal_interpolation_ensi(grid, d_bg, state, 50, num_skipped=skipped)
    torch.cuda.synchronize()
    assert int(skipped.item()) == 0
    assert_bit_exact(whole.cpu().numpy(), host, "EnSI device vs host entry point")
    want = orc.optimal_interpolation_ensi((y, x, None, None), bg, (py, px, None, None), obs, sig, pbg,
                                          B.make_structure(B.BARNES, 10000.0), 50, B.CARTESIAN)
    assert_close(whole.cpu().numpy().reshape(-1, E), want, 2.0, RTOL, "EnSI device entry vs oracle")
    # two ranges cut inside a warm-start block, written into one output; then in place
    parts = torch.full_like(d_bg, -5.0)
    n = ny * nx
    for a, b in ((0, 1000), (1000, n)):
        gd.optimal_interpolation_ensi(grid, d_bg, state, 50, out=parts, first=a, count=b - a)
    assert torch.equal(whole, parts)
    inplace = d_bg.clone()
    gd.optimal_interpolation_ensi(grid, inplace, state, 50, out=inplace)
    assert torch.equal(whole, inplace)
    # a member with a missing value is flagged and left untouched
    bg2 = bg.copy()
    bg2[5, 6, 7] = np.nan
    d_bg2 = torch.from_numpy(bg2).cuda()
    flags2 = gd.valid_members(d_bg2)
    assert flags2.sum() == E - 1 and not flags2[7]
    state2 = gd.EnsembleObservationState(points, obs, sig, pbg, s, member_valid=flags2)
    out2 = gd.optimal_interpolation_ensi(grid, d_bg2, state2, 50).cpu().numpy()
    assert_bit_exact(out2[..., 7], bg2[..., 7], "invalid member untouched")
    host2 = gpp.optimal_interpolation_ensi(grid, bg2, points, obs, sig, pbg, s, 50)
    assert_bit_exact(out2, host2, "EnSI device vs host with an invalid member")


def test_spatial_structure_is_refused_where_unsupported(gpp):
    """a spatially varying structure function handed to an entry point that takes a plain descriptor must raise,
    not analyse with a localization distance of zero."""
    from gridpp_b200 import device as gd
    from gridpp_b200._lib import NotImplementedOnDevice
    y, x = np.meshgrid(np.arange(6, dtype=f32) * 1000, np.arange(7, dtype=f32) * 1000, indexing="ij")
    grid = gpp.Grid(y, x, type=gpp.Cartesian)
    spatial = gpp.BarnesStructure(grid, np.full((6, 7), 2500, f32), np.zeros((6, 7), f32), np.zeros((6, 7), f32))
    points = gpp.Points([1000, 3000], [1000, 4000], type=gpp.Cartesian)
    bg = np.zeros((6, 7, 5), f32)
    with pytest.raises(NotImplementedOnDevice):
        gpp.optimal_interpolation_ensi(grid, bg, points, [1.0, 2.0], [0.5, 0.5], np.zeros((2, 5), f32), spatial, 10)
    with pytest.raises(NotImplementedOnDevice):
        gd.ObservationState(points, [1.0, 2.0], [0.5, 0.5], [0.0, 0.0], spatial)
    with pytest.raises(NotImplementedOnDevice):
        gd.EnsembleObservationState(points, [1.0, 2.0], [0.5, 0.5], np.zeros((2, 5), f32), spatial)
    # ... while optimal_interpolation itself supports it
    out = gpp.optimal_interpolation(grid, bg[..., 0], points, [1.0, 2.0], [0.5, 0.5], [0.0, 0.0], spatial, 10)
    assert np.abs(out).max() > 0


# ------------------------------------------------------------------ full BASELINE.json sizes ------------
def test_full_size_c4_quantile_fast_windows(gpp, orc):
    """Config 4 at full size (8000 x 8000, halfwidth 15, 20 thresholds, quantile 0.5): the filter is local, so any window
    of the full-size result must equal the oracle run on that window plus its halo -- bit for bit -- at the four corners,
    across the middle, and on a block of missing values."""
    rng = np.random.default_rng(1000)
    n, hw = 8000, 15
    f = rng.random((n, n), dtype=f32)
    f[4000:4040, 3000:3040] = np.nan
    f[rng.random((n, n), dtype=f32) < 0.002] = np.nan
    thr = np.linspace(0, 1, 20).astype(f32)
    res = gpp.neighbourhood_quantile_fast(f, 0.5, hw, thr)
    assert res.shape == (n, n) and res.dtype == np.float32
    for (r, c) in ((0, 0), (0, n - 90), (n - 90, 0), (n - 90, n - 90), (3990, 2990), (5000, 1234)):
        r0, r1, c0, c1 = max(0, r - hw), min(n, r + 90 + hw), max(0, c - hw), min(n, c + 90 + hw)
        want = orc.neighbourhood_quantile_fast(f[r0:r1, c0:c1], 0.5, hw, thr)[r - r0:r - r0 + 90, c - c0:c - c0 + 90]
        assert_bit_exact(res[r:r + 90, c:c + 90], want, "C4 window (%d,%d)" % (r, c))
    assert np.isnan(res[4016:4024, 3016:3024]).all()          # windows without a valid value
    inner = res[100:-100, 100:-100]
    ok = ~np.isnan(inner)
    assert 0.3 < float(inner[ok].mean()) < 0.7               # the median of uniform noise


def test_full_size_c5_ensi_subsample(gpp, orc):
    """Config 5 at full size (2500 x 2500 grid, dx 200 m, 20 members, 5000 observations, Barnes 10 km, max_points 50):
    every grid point is independent, so a sample of the full-grid analysis must match the oracle at just those points
    (the Points overload, oi_ensi.cpp:114)."""
    rng = np.random.default_rng(1000)
    n, dx, E, S = 2500, 200.0, 20, 5000
    y, x = np.meshgrid(np.arange(n, dtype=f32) * dx, np.arange(n, dtype=f32) * dx, indexing="ij")
    py, px = (rng.random(S) * n * dx).astype(f32), (rng.random(S) * n * dx).astype(f32)
    bg = rng.standard_normal((n, n, E), dtype=f32)
    bg += rng.standard_normal((n, n, 1), dtype=f32) * 2
    pbg = rng.standard_normal((S, E)).astype(f32)
    obs = rng.standard_normal(S).astype(f32)
    sig = np.full(S, 0.5, f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    out = gpp.optimal_interpolation_ensi(grid, bg, points, obs, sig, pbg, gpp.BarnesStructure(10000), 50)
    assert out.shape == bg.shape and not np.isnan(out).any()
    pick = rng.choice(n * n, 3000, replace=False)
    pick[:4] = [0, n - 1, n * (n - 1), n * n - 1]
    want = orc.optimal_interpolation_ensi((y.ravel()[pick], x.ravel()[pick], None, None), bg.reshape(-1, E)[pick], (py, px, None, None), obs,
                                          sig, pbg, B.make_structure(B.BARNES, 10000.0), 50, B.CARTESIAN)
    assert_close(out.reshape(-1, E)[pick], want, 2.0, RTOL, "C5 subsample")
    assert np.abs(out - bg).max() > 0.1


# ------------------------------------------------------------------ statistics family --
STAT_NAMES = ("mean", "min", "median", "max", "std", "variance", "sum", "count")


def test_statistics_family_golden(gpp):
    """calc_statistic / calc_quantile / interpolate, neighbourhood with Std / Variance / Median, neighbourhood_brute_force and
    neighbourhood_quantile (vec2 and vec3) against the fixture generated from the compiled reference: bit for bit, except
    where a window sum is involved (Std / Variance of neighbourhood() go through the Mean filter: 1e-5)."""
    from util import golden
    g = golden("statistics")
    f, e, rows = g["field"], g["ensemble"], g["rows"]
    code = {n: getattr(gpp, n.capitalize()) for n in STAT_NAMES}
    for key in g.files:
        if key.startswith("nbh_hw"):
            hw, name = int(key[6:].split("__")[0]), key.split("__")[1]
            got = gpp.neighbourhood(f, hw, code[name])
            if name == "median":
                assert_bit_exact(got, g[key], key)
            else:
                # mean2 - mean^2 cancels, and the reference's two means come out of a double summed-area table whose corner
                # differences carry ~1e-11 of rounding noise: a variance that is exactly 0 here can be a tiny negative number
                # there (NaN after the square root). The 1e-5 bar therefore applies to the VARIANCE at the scale of mean2
                # (max |field|^2); a NaN on one side only is a variance within that bar of zero.
                want = g[key]
                tol = 4e-5 * float(np.nanmax(np.abs(f))) ** 2
                v_got, v_want = (got ** 2, want ** 2) if name == "std" else (got, want)
                valid_window = ~np.isnan(gpp.neighbourhood(f, hw, gpp.Mean))
                assert np.isnan(got[~valid_window]).all() and np.isnan(want[~valid_window]).all(), key
                vg = np.where(np.isnan(v_got), 0.0, v_got)[valid_window]
                vw = np.where(np.isnan(v_want), 0.0, v_want)[valid_window]
                assert np.abs(vg - vw).max() <= tol, (key, float(np.abs(vg - vw).max()), tol)
        elif key.startswith("brute_ens_hw") or key.startswith("brute_hw"):
            ens = key.startswith("brute_ens_hw")
            hw, name = int(key.split("hw")[1].split("__")[0]), key.split("__")[1]
            assert_bit_exact(gpp.neighbourhood_brute_force(e if ens else f, hw, code[name]), g[key], key)
        elif key.startswith("quantile_"):
            ens = key.startswith("quantile_ens_hw")
            hw, q = int(key.split("hw")[1].split("__")[0]), float(key.split("__q")[1])
            assert_bit_exact(gpp.neighbourhood_quantile(e if ens else f, q, hw), g[key], key)
        elif key.startswith("rows__q"):
            assert_bit_exact(gpp.calc_quantile(rows, float(key[7:])), g[key], key)
        elif key.startswith("rows__"):
            assert_bit_exact(gpp.calc_statistic(rows, code[key[6:]]), g[key], key)
    assert_bit_exact(gpp.interpolate(g["interp_x"], g["interp_ix"], g["interp_iy"]), g["interp_y"], "interpolate")
    # scalar forms, per-cell quantile levels, error behaviour
    assert gpp.calc_statistic(rows[0], gpp.Mean) == g["rows__mean"][0]
    assert gpp.calc_quantile(rows[3], 0.5) == g["rows__q0.5"][3]
    q2 = np.full((4, 5), 0.25, f32)
    assert_bit_exact(gpp.calc_quantile(rows[:20].reshape(4, 5, -1), q2).ravel(), g["rows__q0.25"][:20], "calc_quantile(vec3, vec2)")
    with pytest.raises(ValueError):
        gpp.calc_quantile(rows[0], 1.1)
    with pytest.raises(ValueError):
        gpp.interpolate(0.5, [0, 1, 2], [0, 1])
    assert np.isnan(gpp.interpolate(0.5, [], []))
    # RandomChoice: any valid value of the row / window is a correct outcome
    picks = gpp.calc_statistic(rows, gpp.RandomChoice)
    for r, p in zip(rows, picks):
        assert (np.isnan(p) and not np.isfinite(r).any()) or p in r
    rc = gpp.neighbourhood(f, 1, gpp.RandomChoice)
    empty = np.isnan(gpp.neighbourhood(f, 1, gpp.Mean))            # windows without a valid value
    assert empty.any() and np.isnan(rc[empty]).all() and not np.isnan(rc[~empty]).any()
    assert rc[0, 0] in f[0:2, 0:2]
    # a large window (961 values) and a window that does not fit the shared-memory stage (31 x 31 x 5 values)
    big = np.random.default_rng(3).random((40, 40)).astype(f32)
    got = gpp.neighbourhood_quantile(big, 0.5, 15)[5, 20]           # rows 0..20, columns 5..35 of the field
    assert abs(got - np.quantile(big[0:21, 5:36].astype(np.float64), 0.5)) < 1e-6
    ens_big = np.random.default_rng(4).random((33, 33, 5)).astype(f32)
    got = gpp.neighbourhood_quantile(ens_big, 1.0, 15)[16, 16]
    assert got == ens_big[1:32, 1:32].max()
    assert gpp.neighbourhood_brute_force(ens_big, 15, gpp.Count)[16, 16] == 31 * 31 * 5


# ------------------------------------------------------------------ consumers of the point index --
def test_gridding_family_golden(gpp):
    """gridding / gridding_nearest / count / distance / fill / fill_missing / doping_square / doping_circle against the fixture
    generated from the compiled reference. Integer-valued and copied results are bit-exact; sums and the great-circle
    distance (double trigonometry on the device vs glibc) within 1e-5."""
    from util import golden
    import test_oracle
    g = golden("gridding")
    stat = {n: getattr(gpp, n.capitalize()) for n in STAT_NAMES}
    exact_stats = ("min", "max", "median", "count")
    n = 0
    for key, tag, ctype, a, sets, func, spec in test_oracle._gridding_cases(g):
        t = gpp.Cartesian if ctype == B.CARTESIAN else gpp.Geodetic
        obj = dict(grid=gpp.Grid(a["glats"], a["glons"], a["gelevs"], type=t), points=gpp.Points(a["plats"], a["plons"], a["pelevs"], type=t),
                   ogrid=gpp.Grid(*sets["ogrid"], type=t), opoints=gpp.Points(*sets["opoints"], type=t))
        radius = float(a["radius"])
        want = g[key]
        if func.startswith("gridding"):
            nearest = "nearest" in func
            out = obj["grid"] if func.endswith("grid") else obj["opoints"]
            name = spec.split("_mn")[0]
            mn = int(spec.split("_mn")[1]) if "_mn" in spec else (0 if nearest else 1)
            got = gpp.gridding_nearest(out, obj["points"], a["values"], mn, stat[name]) if nearest else \
                gpp.gridding(out, obj["points"], a["values"], radius, mn, stat[name])
            if name in exact_stats or name in ("mean", "sum", "std", "variance"):
                assert_bit_exact(got, want, key)        # float accumulation in the reference's order: the same bits
        elif func == "count":
            i, o = spec.split("_")
            assert_bit_exact(gpp.count(obj[i], obj[o], radius), want, key)
        elif func == "distance":
            i, o, num = spec.split("_")
            got = gpp.distance(obj[i], obj[o], int(num[1:]))
            if ctype == B.CARTESIAN:
                assert_bit_exact(got, want, key)
            else:
                assert_close(got, want, 1000.0, RTOL, key)
        elif func == "fill":
            assert_bit_exact(gpp.fill(obj["grid"], a["field"], obj["points"], a["radii"], -7.5, bool(int(spec[-1]))), want, key)
        else:
            med = np.nan if spec == "nocheck" else 150.0
            if func == "doping_square":
                got = gpp.doping_square(obj["grid"], a["field"], obj["points"], a["values"], a["halfwidth"], med)
            else:
                got = gpp.doping_circle(obj["grid"], a["field"], obj["points"], a["values"], a["radii"], med)
            assert_bit_exact(got, want, key)
        n += 1
    assert n >= 100
    assert_bit_exact(gpp.fill_missing(g["fill_missing__in"]), g["fill_missing__out"], "fill_missing")
    # argument checks (gridding.cpp:7-12, fill.cpp:7-14, doping.cpp:6-28)
    grid, points = obj["grid"], obj["points"]
    with pytest.raises(ValueError):
        gpp.gridding(grid, points, a["values"][:-1], radius, 0, gpp.Mean)
    with pytest.raises(ValueError):
        gpp.gridding(grid, points, a["values"], -1, 0, gpp.Mean)
    with pytest.raises(ValueError):
        gpp.gridding_nearest(grid, points, a["values"], -1, gpp.Mean)
    with pytest.raises(ValueError):
        gpp.fill(grid, a["field"], points, -a["radii"] - 1, 0, False)
    with pytest.raises(ValueError):
        gpp.doping_circle(grid, a["field"], points, a["values"], a["radii"], -1.0)
    with pytest.raises(ValueError):
        gpp.distance(gpp.Points([0], [0]), gpp.Points([0], [0], type=gpp.Cartesian))


def test_gridding_large_random_vs_oracle(gpp, orc):
    """A denser case than the fixture (many neighbours per node, chunked pair lists): 300 x 300 grid, 20 000 points."""
    rng = np.random.default_rng(5)
    ny, nx, S = 300, 300, 20000
    y, x = np.meshgrid(np.arange(ny, dtype=f32) * 500, np.arange(nx, dtype=f32) * 500, indexing="ij")
    py, px = rng.uniform(0, ny * 500, S).astype(f32), rng.uniform(0, nx * 500, S).astype(f32)
    v = rng.normal(size=S).astype(f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    for st_g, st_o in ((gpp.Mean, B.MEAN), (gpp.Median, B.MEDIAN), (gpp.Count, B.COUNT)):
        got = gpp.gridding(grid, points, v, 4000.0, 5, st_g)
        want = orc.gridding((y, x), (py, px), v, 4000.0, 5, st_o, B.CARTESIAN)
        assert_bit_exact(got, want, "gridding statistic %d" % st_o)
    assert_bit_exact(gpp.count(points, grid, 2500.0), orc.count((py, px), (y, x), 2500.0, B.CARTESIAN), "count")
    assert_bit_exact(gpp.gridding_nearest(grid, points, v, 0, gpp.Max), orc.gridding((y, x), (py, px), v, 0, 0, B.MAX, B.CARTESIAN, nearest=True),
                     "gridding_nearest")


# ------------------------------------------------------------------ ensi_multi / staticcorr_points --
def _multi_sets(gpp, a, ctype):
    t = gpp.Cartesian if ctype == B.CARTESIAN else gpp.Geodetic
    return gpp.Points(a["by"], a["bx"], a["be"], a["bf"], t), gpp.Points(a["py"], a["px"], a["pe"], a["pf"], t)


def _multi_call(gpp, kind, bp, op, a, bg, pbg, s, mp, extr):
    if kind == "ebesc":
        return gpp.optimal_interpolation_ensi_multi_ebesc(bp, a["bratios"], bg, op, a["pobs2"], a["pratios"], pbg, s, mp, bool(extr))
    if kind == "ebe":
        return gpp.optimal_interpolation_ensi_multi_ebe(bp, a["bratios"], bg, a["background_corr"], op, a["pobs2"], a["pratios"], pbg,
                                                        a["pbackground_corr"], s, mp, bool(extr))
    return gpp.optimal_interpolation_ensi_multi_utem(bp, a["bratios"], bg, a["background_corr"], op, a["pobs1"], a["pratios"], pbg,
                                                     a["pbackground_corr"], s, mp, bool(extr))


def test_ensi_multi_and_staticcorr_golden(gpp):
    """optimal_interpolation_ensi_multi_{ebe,ebesc,utem} and staticcorr_points against the fixture written by the compiled
    reference (tests/golden/make_golden_ensi_multi.py): Cartesian Barnes with elevation and land-area terms, Geodetic Cressman;
    unlimited and max_points = 12; with and without the anti-extrapolation filter; two trailing invalid members.
    staticcorr_points is pure float arithmetic: bit for bit. The analyses go through a k x k solve (elimination here, an explicit
    inverse in the reference) or an eigen-decomposition: the 1e-5 bar of SURVEY.md 8(d). The clamp is a branch on the
    increment, so a value within rounding of a bound may take the other branch: at most a handful of outliers are tolerated and
    their size is bounded separately."""
    from test_oracle import _ensi_multi_cases
    from util import golden
    g = golden("ensi_multi")
    n_runs = 0
    for tag, ctype, spec, a, runs in _ensi_multi_cases(g):
        bp, op = _multi_sets(gpp, a, ctype)
        s = (gpp.BarnesStructure if spec[0] == B.BARNES else gpp.CressmanStructure)(*spec[1:])
        for mp in (0, 12):
            assert_bit_exact(gpp.staticcorr_points(bp, op, s, mp), a["staticcorr_mp%d" % mp], "%s staticcorr_points mp=%d" % (tag, mp))
        for key, kind, mp, extr, bg, pbg in runs:
            what = "%s %s" % (tag, key)
            if kind == "utem" and mp == 0 and (a["staticcorr_mp0"] != 0).sum(1).max() > 64:
                with pytest.raises(RuntimeError, match="at most 64 observations"):   # the EnSI kernel holds at most 64 observations per point
                    _multi_call(gpp, kind, bp, op, a, bg, pbg, s, mp, extr)
                continue
            got = _multi_call(gpp, kind, bp, op, a, bg, pbg, s, mp, extr)
            want = a[key]
            assert got.shape == want.shape and got.dtype == want.dtype
            assert_close(got, want, 1.0, RTOL, what, allow_outliers=0 if extr else 4)
            assert np.nanmax(np.abs(got - want)) < 1e-2, what
            untouched = (want == bg) | np.isnan(bg)
            assert np.array_equal(got[untouched], bg[untouched], equal_nan=True), what + ": members / points the reference leaves alone"
            n_runs += 1
    assert n_runs >= 24
    # an invalid member in the middle: the reference's innovation matrix is indexed out of bounds (Armadillo throws)
    tag, ctype, spec, a, runs = next(_ensi_multi_cases(g))
    bp, op = _multi_sets(gpp, a, ctype)
    bg = a["background"].copy()
    bg[0, 2] = np.nan
    with pytest.raises(RuntimeError):
        gpp.optimal_interpolation_ensi_multi_ebesc(bp, a["bratios"], bg, op, a["pobs2"], a["pratios"], a["pbackground"], gpp.BarnesStructure(*spec[1:]), 12)
    out = gpp.optimal_interpolation_ensi_multi_utem(bp, a["bratios"], bg, a["background_corr"], op, a["pobs1"], a["pratios"], a["pbackground"],
                                                    a["pbackground_corr"], gpp.BarnesStructure(*spec[1:]), 12)
    assert np.isnan(out[0, 2]) and np.array_equal(out[1:, 2], bg[1:, 2]) and not np.isnan(np.delete(out, 2, axis=1)).any()


def test_ensi_multi_grid_overloads_vs_oracle(gpp, orc):
    """The Grid overloads (oi_ensi_multi.cpp:34-327 flatten the grid) on a 60 x 70 grid against the oracle's Points form, and the
    argument checks of the reference."""
    rng = np.random.default_rng(5)
    ny, nx, S, E, dx, mp = 60, 70, 120, 10, 1000.0, 15
    y, x = np.meshgrid(np.arange(ny, dtype=f32) * dx, np.arange(nx, dtype=f32) * dx, indexing="ij")
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    bg = (rng.standard_normal((ny, nx, E)) + 2 * np.sin(y / 9000.0)[:, :, None]).astype(f32)
    bgc = (rng.standard_normal((ny, nx, E)) + np.cos(x / 7000.0)[:, :, None]).astype(f32)
    pbg, pbgc = rng.standard_normal((S, E)).astype(f32), rng.standard_normal((S, E)).astype(f32)
    pobs2 = (pbg + 0.8 + 0.3 * rng.standard_normal((S, E))).astype(f32)
    pobs1 = pobs2[:, 0].copy()
    pratios, bratios = rng.uniform(0.1, 0.5, S).astype(f32), rng.uniform(0.8, 1.2, (ny, nx)).astype(f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    s, so = gpp.BarnesStructure(7000), B.make_structure(B.BARNES, 7000.0)
    bp, op = (y.ravel(), x.ravel(), None, None), (py, px, None, None)
    for extr in (True, False):
        got = gpp.optimal_interpolation_ensi_multi_ebesc(grid, bratios, bg, points, pobs2, pratios, pbg, s, mp, extr)
        want = orc.ensi_multi("ebesc", bp, bratios, bg, None, op, pobs2, pratios, pbg, None, so, mp, B.CARTESIAN, extr)
        assert_close(got.reshape(-1, E), want, 1.0, RTOL, "ebesc grid", allow_outliers=0 if extr else 4)
        got = gpp.optimal_interpolation_ensi_multi_ebe(grid, bratios, bg, bgc, points, pobs2, pratios, pbg, pbgc, s, mp, extr)
        want = orc.ensi_multi("ebe", bp, bratios, bg, bgc, op, pobs2, pratios, pbg, pbgc, so, mp, B.CARTESIAN, extr)
        assert_close(got.reshape(-1, E), want, 1.0, RTOL, "ebe grid", allow_outliers=0 if extr else 4)
        got = gpp.optimal_interpolation_ensi_multi_utem(grid, bratios, bg, bgc, points, pobs1, pratios, pbg, pbgc, s, mp, extr)
        want = orc.ensi_multi("utem", bp, bratios, bg, bgc, op, pobs1, pratios, pbg, pbgc, so, mp, B.CARTESIAN, extr)
        assert_close(got.reshape(-1, E), want, 1.0, RTOL, "utem grid", allow_outliers=0 if extr else 4)
        assert got.shape == bg.shape and np.abs(got - bg).max() > 0.05
    with pytest.raises(ValueError):
        gpp.optimal_interpolation_ensi_multi_ebesc(grid, bratios, bg, points, pobs2, pratios, pbg, s, -1)
    with pytest.raises(ValueError):
        gpp.optimal_interpolation_ensi_multi_ebesc(grid, bratios[:-1], bg, points, pobs2, pratios, pbg, s, mp)
    with pytest.raises(ValueError):
        gpp.optimal_interpolation_ensi_multi_utem(grid, bratios, bg, bgc, points, pobs2, pratios, pbg, pbgc, s, mp)   # pobs must be (S,)
    with pytest.raises(ValueError):
        gpp.staticcorr_points(points, gpp.Points(py, px), s, 3)   # coordinate types differ
    none = gpp.Points([], [], type=gpp.Cartesian)
    assert np.array_equal(gpp.optimal_interpolation_ensi_multi_ebe(grid, bratios, bg, bgc, none, np.zeros((0, E), f32), [], np.zeros((0, E), f32),
                                                                   np.zeros((0, E), f32), s, mp), bg)
    assert gpp.staticcorr_points(points, none, s, 0).shape == (S, 0)


# ------------------------------------------------------------------ device-resident chaining --
def test_device_resident_chain_equals_host_calls(gpp):
    """the analysis stays in HBM and is post-processed there -- OI -> neighbourhood mean -> quantile_fast, all on
    the current stream, no host round trip -- and equals, bit for bit, the same three calls made one by one through the host
    API (numpy in, numpy out)."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(11)
    ny, nx, dx = 300, 320, 500.0
    y, x, py, px, bg, pbg, obs, ratios = _c3_like(rng, ny, nx, dx, density=0.02)
    grid, points, s = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(8000)
    thr = np.linspace(-8, 8, 17).astype(f32)
    state = gd.ObservationState(points, obs, ratios, pbg, s)
    d_bg = torch.from_numpy(bg.ravel()).cuda()
    d_an = gd.optimal_interpolation(grid, d_bg, state, 25).reshape(ny, nx)
    d_mean = gd.neighbourhood(d_an, 5, gpp.Mean)
    d_q = gd.neighbourhood_quantile_fast(d_mean, 0.9, 7, thr)
    torch.cuda.synchronize()
    an = gpp.optimal_interpolation(grid, bg, points, obs, ratios, pbg, s, 25)
    mean = gpp.neighbourhood(an, 5, gpp.Mean)
    q = gpp.neighbourhood_quantile_fast(mean, 0.9, 7, thr)
    assert_bit_exact(d_an.cpu().numpy(), an, "analysis")
    assert_bit_exact(d_mean.cpu().numpy(), mean, "mean of the analysis")
    assert_bit_exact(d_q.cpu().numpy(), q, "quantile of the mean")
    assert np.isfinite(q).all() and np.ptp(q) > 0


# ------------------------------------------------------------------ neighbourhood_search / calc_gradient --
def test_window_filters_golden(gpp, orc):
    """neighbourhood_search and calc_gradient against the fixture written by the compiled reference
    (tests/golden/make_golden_window_filters.py). neighbourhood_search and the MinMax gradient walk the window in the reference's
    order with the reference's float operations: bit for bit. The LinearRegression gradient is built on five neighbourhood means
    (within 1e-6 of the reference's double summed-area tables) and divides a difference of moments by a variance, which amplifies
    that by |mean|^2 / variance: compared at 1e-5 of the gradient's scale where the variance is not dominated by cancellation,
    and on a larger random case against the oracle."""
    from util import golden
    g = golden("window_filters")
    a, s, ap = g["search__array"], g["search__search"], g["search__apply"]
    for i, (hw, lo, hi, delta) in enumerate(g["search__cases"]):
        assert_bit_exact(gpp.neighbourhood_search(a, s, int(hw), lo, hi, delta), g["search__case%d" % i], "search %d" % i)
        assert_bit_exact(gpp.neighbourhood_search(a, s, int(hw), lo, hi, delta, ap), g["search__case%d_apply" % i], "search %d + apply" % i)
    base, values = g["gradient__base"], g["gradient__values"]
    for i, (hw, num_min, min_range, default) in enumerate(g["gradient__cases"]):
        hw, num_min = int(hw), int(num_min)
        assert_bit_exact(gpp.calc_gradient(base, values, gpp.MinMax, hw, num_min, min_range, default), g["gradient__minmax_case%d" % i], "MinMax %d" % i)
        got = gpp.calc_gradient(base, values, gpp.LinearRegression, hw, num_min, min_range, default)
        want = g["gradient__regression_case%d" % i]
        # the elevations are ~300 +- 200: variance / mean^2 ~ 0.4, no cancellation to speak of
        decided_same = (got == default) == (want == default)
        assert decided_same.mean() > 0.999, "regression %d: default-gradient decisions differ at %d points" % (i, (~decided_same).sum())
        assert_close(got[decided_same], want[decided_same], 0.01, 1e-4, "regression %d" % i)
    # a larger random case against the oracle
    rng = np.random.default_rng(3)
    b = rng.standard_normal((300, 280)).astype(f32)
    v = (2.5 * b + 0.3 * rng.standard_normal(b.shape)).astype(f32)
    b[rng.uniform(size=b.shape) < 0.02] = np.nan
    assert_bit_exact(gpp.calc_gradient(b, v, gpp.MinMax, 3), orc.calc_gradient(b, v, 0, 3), "MinMax random")
    assert_close(gpp.calc_gradient(b, v, gpp.LinearRegression, 3), orc.calc_gradient(b, v, 10, 3), 1.0, 1e-4, "regression random")
    sa = rng.uniform(0, 1, b.shape).astype(f32)
    assert_bit_exact(gpp.neighbourhood_search(v, sa, 2, 0.9, 1.0, 0.05), orc.neighbourhood_search(v, sa, 2, 0.9, 1.0, 0.05), "search random")
    with pytest.raises(ValueError):
        gpp.neighbourhood_search(v, sa, 2, 1.0, 0.9, 0.05)
    with pytest.raises(ValueError):
        gpp.neighbourhood_search(v, sa[:-1], 2, 0.9, 1.0, 0.05)
    with pytest.raises(ValueError):
        gpp.calc_gradient(b, v, gpp.MinMax, 0)


def test_closest_neighbours_any_number(gpp, orc):
    """KDTree::get_closest_neighbours (kdtree.cpp:82-103) for more neighbours than the register kernel holds (16): the general
    kernel keeps the sorted list in global memory. Bit-equal to the oracle incl. duplicates (ties -> lowest index), more
    neighbours than points, include_match = False; and gridpp::distance with num > 16."""
    rng = np.random.default_rng(21)
    for t, ctype, span in ((gpp.Cartesian, B.CARTESIAN, 50000.0), (gpp.Geodetic, B.GEODETIC, 3.0)):
        n = 3000
        lats, lons = (rng.uniform(0, span, n)).astype(f32), (rng.uniform(0, span, n)).astype(f32)
        lats[100:110], lons[100:110] = lats[99], lons[99]                       # duplicates
        ql, qo = rng.uniform(-0.1 * span, 1.1 * span, 200).astype(f32), rng.uniform(-0.1 * span, 1.1 * span, 200).astype(f32)
        ql[:50], qo[:50] = lats[:50], lons[:50]                                 # queries on top of points
        p = gpp.Points(lats, lons, type=t)
        for num, match in ((17, True), (40, True), (40, False), (333, True)):
            assert_bit_exact(p._set.closest(ql, qo, num, match), orc.points_closest(lats, lons, ctype, ql, qo, num, match), "closest %d" % num)
        small = gpp.Points(lats[:25], lons[:25], type=t)
        got = small._set.closest(ql[:20], qo[:20], 60, True)
        assert got.shape == (20, 60) and (got[:, :25] >= 0).all() and (got[:, 25:] == -1).all()
        assert len(small.get_closest_neighbours(float(ql[0]), float(qo[0]), 60)) == 25
        assert_bit_exact(gpp.distance(p, gpp.Points(ql, qo, type=t), 30), orc.distance((lats, lons), (ql, qo), 30, ctype), "distance num=30")


def test_ensi_multi_many_members_and_limits(gpp, orc):
    """ebe / ebesc take any number of members (loops over members, 40 here) and up to 128 observations per point; utem shares the
    EnSI kernel's limits (32 valid members, 64 observations) and says so instead of computing something else."""
    rng = np.random.default_rng(9)
    nB, S, E = 500, 300, 40
    by, bx = rng.uniform(0, 30000, nB).astype(f32), rng.uniform(0, 30000, nB).astype(f32)
    py, px = rng.uniform(0, 30000, S).astype(f32), rng.uniform(0, 30000, S).astype(f32)
    bg, bgc = rng.standard_normal((nB, E)).astype(f32), rng.standard_normal((nB, E)).astype(f32)
    pbg, pbgc = rng.standard_normal((S, E)).astype(f32), rng.standard_normal((S, E)).astype(f32)
    pobs2 = (pbg + 0.5 + 0.2 * rng.standard_normal((S, E))).astype(f32)
    pratios, bratios = rng.uniform(0.1, 0.4, S).astype(f32), np.ones(nB, f32)
    bp, op = gpp.Points(by, bx, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    s, so = gpp.BarnesStructure(2500), B.make_structure(B.BARNES, 2500.0)
    sets = ((by, bx, None, None), (py, px, None, None))
    for mp in (20, 0):   # 0: unlimited, ~60 observations in reach of a point here, up to ~100
        got = gpp.optimal_interpolation_ensi_multi_ebe(bp, bratios, bg, bgc, op, pobs2, pratios, pbg, pbgc, s, mp)
        want = orc.ensi_multi("ebe", sets[0], bratios, bg, bgc, sets[1], pobs2, pratios, pbg, pbgc, so, mp, B.CARTESIAN, True)
        assert_close(got, want, 1.0, RTOL, "ebe E=40 mp=%d" % mp)
        got = gpp.optimal_interpolation_ensi_multi_ebesc(bp, bratios, bg, op, pobs2, pratios, pbg, s, mp, False)
        want = orc.ensi_multi("ebesc", sets[0], bratios, bg, None, sets[1], pobs2, pratios, pbg, None, so, mp, B.CARTESIAN, False)
        assert_close(got, want, 1.0, RTOL, "ebesc E=40 mp=%d" % mp, allow_outliers=4)
    with pytest.raises(RuntimeError, match="at most 32 valid ensemble members"):
        gpp.optimal_interpolation_ensi_multi_utem(bp, bratios, bg, bgc, op, pobs2[:, 0], pratios, pbg, pbgc, s, 20)
    wide = gpp.BarnesStructure(9000)   # several hundred observations in reach
    with pytest.raises(RuntimeError, match="at most 128 observations"):
        gpp.optimal_interpolation_ensi_multi_ebesc(bp, bratios, bg, op, pobs2, pratios, pbg, wide, 0)
    assert gpp.staticcorr_points(bp, op, wide, 0).shape == (nB, S)
    with pytest.raises(RuntimeError, match="at most 128"):
        gpp.staticcorr_points(bp, op, wide, 200)


def test_ensi_multi_pipelined_path_subsample(gpp, orc):
    """Fields of 2^18 points or more go through the host pipeline (member scan on host threads, blocks uploaded / analysed /
    returned on alternating streams). Every point is independent, so a sample of a 520 x 520 analysis must equal the oracle's
    Points form on just those points; an invalid trailing member must come back untouched."""
    rng = np.random.default_rng(13)
    n, dx, E, S, mp = 520, 400.0, 6, 2500, 12
    y, x = np.meshgrid(np.arange(n, dtype=f32) * dx, np.arange(n, dtype=f32) * dx, indexing="ij")
    py, px = rng.uniform(0, n * dx, S).astype(f32), rng.uniform(0, n * dx, S).astype(f32)
    bg = (rng.standard_normal((n, n, E)) + 2 * np.sin(y / 30000.0)[:, :, None]).astype(f32)
    bgc = rng.standard_normal((n, n, E)).astype(f32)
    bg[17, 300, E - 1] = np.nan                                   # the last member is invalid somewhere: left alone everywhere
    pbg, pbgc = rng.standard_normal((S, E)).astype(f32), rng.standard_normal((S, E)).astype(f32)
    pobs2 = (pbg + 0.6 + 0.3 * rng.standard_normal((S, E))).astype(f32)
    pratios, bratios = rng.uniform(0.1, 0.5, S).astype(f32), rng.uniform(0.8, 1.2, (n, n)).astype(f32)
    grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    s, so = gpp.BarnesStructure(6000), B.make_structure(B.BARNES, 6000.0)
    pick = rng.choice(n * n, 2500, replace=False)
    pick[:4] = [0, n - 1, n * (n - 1), n * n - 1]
    bp = (y.ravel()[pick], x.ravel()[pick], None, None)
    flat = lambda a: a.reshape(n * n, -1)[pick]
    for kind in ("ebesc", "ebe", "utem"):
        if kind == "ebesc":
            got = gpp.optimal_interpolation_ensi_multi_ebesc(grid, bratios, bg, points, pobs2, pratios, pbg, s, mp, False)
        elif kind == "ebe":
            got = gpp.optimal_interpolation_ensi_multi_ebe(grid, bratios, bg, bgc, points, pobs2, pratios, pbg, pbgc, s, mp, False)
        else:
            got = gpp.optimal_interpolation_ensi_multi_utem(grid, bratios, bg, bgc, points, pobs2[:, 0], pratios, pbg, pbgc, s, mp, False)
        # the oracle sees only the sample: give it the same member validity by keeping the invalid value in the sample
        sample_bg = flat(bg).copy()
        sample_bg[5, E - 1] = np.nan
        corr = kind != "ebesc"
        want = orc.ensi_multi(kind, bp, bratios.ravel()[pick], sample_bg, flat(bgc) if corr else None, (py, px, None, None),
                              pobs2[:, 0] if kind == "utem" else pobs2, pratios, pbg, pbgc if corr else None, so, mp, B.CARTESIAN, False)
        g = flat(got).copy()
        g[5, E - 1] = np.nan
        assert_close(g[:, :E - 1], want[:, :E - 1], 1.0, RTOL, kind + " pipelined", allow_outliers=4)
        assert np.array_equal(got[..., E - 1], bg[..., E - 1], equal_nan=True), kind + ": the invalid member must be untouched"
        assert np.abs(got[..., :E - 1] - bg[..., :E - 1]).max() > 0.1


def test_next_rows_edge_cases(gpp, orc):
    """Degenerate shapes of the SURVEY 8(f) rows against the oracle: a single member (the standardisation divides by
    sqrt(E - 1) = 0 and falls back to zeros), a single observation, max_points = 1, observations that are all invalid, a point
    with no observation in reach, a 1 x 1 field for the window filters."""
    rng = np.random.default_rng(31)
    nB, S = 60, 25
    by, bx = rng.uniform(0, 10000, nB).astype(f32), rng.uniform(0, 10000, nB).astype(f32)
    by[0], bx[0] = 90000.0, 90000.0                                            # nothing in reach: stays at the background
    py, px = rng.uniform(0, 10000, S).astype(f32), rng.uniform(0, 10000, S).astype(f32)
    bp, op = gpp.Points(by, bx, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
    s, so = gpp.BarnesStructure(3000), B.make_structure(B.BARNES, 3000.0)
    sets = ((by, bx, None, None), (py, px, None, None))
    pr, br = rng.uniform(0.1, 0.5, S).astype(f32), rng.uniform(0.8, 1.2, nB).astype(f32)
    for E in (1, 2, 5):
        bg, bgc = rng.standard_normal((nB, E)).astype(f32), rng.standard_normal((nB, E)).astype(f32)
        pbg, pbgc = rng.standard_normal((S, E)).astype(f32), rng.standard_normal((S, E)).astype(f32)
        pobs2 = (pbg + 0.5).astype(f32)
        for mp in (1, 7, 0):
            what = "E=%d mp=%d" % (E, mp)
            got = gpp.optimal_interpolation_ensi_multi_ebesc(bp, br, bg, op, pobs2, pr, pbg, s, mp, False)
            want = orc.ensi_multi("ebesc", sets[0], br, bg, None, sets[1], pobs2, pr, pbg, None, so, mp, B.CARTESIAN, False)
            assert_close(got, want, 1.0, RTOL, "ebesc " + what, allow_outliers=2)
            assert np.array_equal(got[0], bg[0])
            got = gpp.optimal_interpolation_ensi_multi_ebe(bp, br, bg, bgc, op, pobs2, pr, pbg, pbgc, s, mp, True)
            want = orc.ensi_multi("ebe", sets[0], br, bg, bgc, sets[1], pobs2, pr, pbg, pbgc, so, mp, B.CARTESIAN, True)
            assert_close(got, want, 1.0, RTOL, "ebe " + what)
            if E > 1:
                got = gpp.optimal_interpolation_ensi_multi_utem(bp, br, bg, bgc, op, pobs2[:, 0], pr, pbg, pbgc, s, mp, True)
                want = orc.ensi_multi("utem", sets[0], br, bg, bgc, sets[1], pobs2[:, 0], pr, pbg, pbgc, so, mp, B.CARTESIAN, True)
                assert_close(got, want, 1.0, RTOL, "utem " + what)
        none_valid = np.full((S, E), np.nan, f32)
        assert np.array_equal(gpp.optimal_interpolation_ensi_multi_ebesc(bp, br, bg, op, none_valid, pr, pbg, s, 5), bg)
        one = gpp.Points(py[:1], px[:1], type=gpp.Cartesian)
        got = gpp.optimal_interpolation_ensi_multi_ebesc(bp, br, bg, one, pobs2[:1], pr[:1], pbg[:1], s, 5)
        want = orc.ensi_multi("ebesc", sets[0], br, bg, None, (py[:1], px[:1], None, None), pobs2[:1], pr[:1], pbg[:1], None, so, 5, B.CARTESIAN, True)
        assert_close(got, want, 1.0, RTOL, "one observation, E=%d" % E)
    assert_bit_exact(gpp.staticcorr_points(bp, op, s, 1), orc.staticcorr_points(sets[0], sets[1], so, 1, B.CARTESIAN), "staticcorr mp=1")
    assert_bit_exact(gpp.staticcorr_points(bp, gpp.Points(py[:1], px[:1], type=gpp.Cartesian), s, 0),
                     orc.staticcorr_points(sets[0], (py[:1], px[:1], None, None), so, 0, B.CARTESIAN), "staticcorr one knot")
    one_cell = np.array([[2.5]], f32)
    assert_bit_exact(gpp.neighbourhood_search(one_cell, one_cell, 3, 0.0, 1.0, 0.1), orc.neighbourhood_search(one_cell, one_cell, 3, 0.0, 1.0, 0.1), "search 1x1")
    for gt in (gpp.MinMax, gpp.LinearRegression):
        assert_bit_exact(gpp.calc_gradient(one_cell, one_cell, gt, 2, 0, 0.0, -3.0), orc.calc_gradient(one_cell, one_cell, int(gt), 2, 0, 0.0, -3.0), "gradient 1x1")
    f = rng.standard_normal((5, 300)).astype(f32)
    assert_bit_exact(gpp.neighbourhood_search(f, f, 0, -0.5, 0.5, 0.0), orc.neighbourhood_search(f, f, 0, -0.5, 0.5, 0.0), "search hw=0")


def test_oi_cholesky_path_traversals(gpp, orc):
    """30 < max_points <= 128 (and unlimited) take the Cholesky kernel: 4 x 8 tiles on whole rows of a grid, 32 consecutive points
    otherwise, a per-warp cache of solved systems. The increment of a point is the same dot product whichever way it is
    reached: Grid traversal == Points traversal == arbitrary ranges, bit for bit; a sample against the oracle."""
    import torch
    from gridpp_b200 import device as gd
    rng = np.random.default_rng(4)
    ny, nx, dx = 203, 301, 250.0          # neither a multiple of 4 rows nor of 8 columns
    y, x, py, px, bg, pbg, obs, ratios = _c3_like(rng, ny, nx, dx)
    bg[rng.uniform(size=bg.shape) < 0.002] = np.nan
    grid, pts = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(y.ravel(), x.ravel(), type=gpp.Cartesian)
    points, s = gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    state = gd.ObservationState(points, obs, ratios, pbg, s)
    d_bg = torch.from_numpy(bg.ravel()).cuda()
    for mp in (50, 0):
        whole = gd.optimal_interpolation(grid, d_bg, state, mp)
        again = gd.optimal_interpolation(grid, d_bg, state, mp)
        flat = gd.optimal_interpolation(pts, d_bg, state, mp)
        parts = torch.full_like(d_bg, float("nan"))
        cuts = [0, 301 * 7, 301 * 7 + 13, 40000, ny * nx]
        for a, b in zip(cuts[:-1], cuts[1:]):
            gd.optimal_interpolation(grid, d_bg, state, mp, out=parts, first=a, count=b - a)
        torch.cuda.synchronize()
        assert torch.equal(whole.isnan(), torch.from_numpy(np.isnan(bg.ravel())).cuda())
        for other, what in ((again, "second launch"), (flat, "Points traversal"), (parts, "ranges")):
            assert torch.equal(torch.nan_to_num(whole), torch.nan_to_num(other)), "max_points %d: %s" % (mp, what)
        pick = rng.choice(ny * nx, 1500, replace=False)
        want = orc.optimal_interpolation((y.ravel()[pick], x.ravel()[pick], None, None), bg.ravel()[pick], (py, px, None, None), obs, ratios, pbg,
                                         B.make_structure(B.BARNES, 10000.0), mp, B.CARTESIAN)
        assert_close(whole.cpu().numpy()[pick], want, 1.0, RTOL, "Cholesky path, max_points %d" % mp)
