"""Generates the golden fixtures of this directory from the UNMODIFIED reference sources.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

It calls oracle/_ref/libgridpp_ref.so (metno/gridpp src/api/*.cpp compiled against the shim headers in
oracle/shims) on small seeded inputs and stores inputs + outputs as .npz. The fixtures travel with the repo;
/root/reference does not exist on the GPU box. Seeds follow the reference's own convention (1000,
tests/benchmark.py:21).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def struct_spec(kind, *args, **kw):
    """A picklable description of a structure function, turned into a descriptor by tests/util.py."""
    return np.array([repr((kind, args, kw))])


def build(spec):
    kind, args, kw = eval(str(spec[0]))
    if kind == "single":
        return B.make_structure(*args, **kw)
    if kind == "cv":
        return B.cross_validation(build(np.array([repr(args[0])])), args[1])
    if kind == "multiple":
        return B.multiple_structure(*[build(np.array([repr(a)])) for a in args])
    raise ValueError(kind)


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)   # neighbourhood_quantile_fast with a quantile field races under OpenMP (neighbourhood.cpp:365-373)
    rng = np.random.default_rng(1000)
    f32 = np.float32

    # ---- OI: config 1 of BASELINE.json (README example shape, Geodetic, 10 obs, max_points 10), scaled to 40x40
    ny = nx = 40
    lats, lons = np.meshgrid(np.linspace(0, 1, ny), np.linspace(0, 1, nx), indexing="ij")
    bg = rng.normal(size=(ny, nx)).astype(f32) * 3
    plats, plons = rng.uniform(0, 1, 10).astype(f32), rng.uniform(0, 1, 10).astype(f32)
    obs = (rng.normal(size=10) / 2).astype(f32)
    idx = ref.points_nearest(lats, lons, B.GEODETIC, plats, plons)
    pbg = bg.ravel()[idx]
    spec = ("single", (B.BARNES, 10000.0), {})
    s = build(np.array([repr(spec)]))
    out, var = ref.optimal_interpolation((lats, lons, None, None), bg, (plats, plons, None, None), obs, np.full(10, 0.5, f32),
                                         pbg, s, 10, B.GEODETIC, want_variance=True)
    np.savez_compressed(os.path.join(HERE, "oi_c1_geodetic.npz"), lats=lats.astype(f32), lons=lons.astype(f32), background=bg,
                        plats=plats, plons=plons, pobs=obs, pratios=np.full(10, 0.5, f32), pbackground=pbg,
                        structure=np.array([repr(spec)]), max_points=10, ctype=B.GEODETIC, nearest_index=idx,
                        analysis=out.reshape(ny, nx), analysis_variance=var.reshape(ny, nx))

    # ---- OI: config 3 density (Cartesian, dx 250 m ... here a 48x48 window, ~42 candidates per point, max_points 30)
    ny = nx = 48
    dx = 2500.0
    yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    S = 150
    py, px = rng.uniform(-2e4, ny * dx + 2e4, S).astype(f32), rng.uniform(-2e4, nx * dx + 2e4, S).astype(f32)
    bg = rng.normal(size=(ny, nx)).astype(f32)
    bg[3, 4] = np.nan
    pbg = rng.normal(size=S).astype(f32)
    obs = (pbg + rng.normal(size=S) * 0.5).astype(f32)
    obs[5] = np.nan
    pbg[9] = np.nan
    ratios = rng.uniform(0.1, 1.0, S).astype(f32)
    cases = {
        "barnes_mp30": (("single", (B.BARNES, 10000.0), {}), 30, True, False),
        "barnes_mp5_clamp": (("single", (B.BARNES, 10000.0), {}), 5, False, False),
        "barnes_elev_unlimited": (("single", (B.BARNES, 10000.0, 200.0, 0.5), {}), 0, True, True),
        "cressman_elev_mp12": (("single", (B.CRESSMAN, 20000.0, 300.0), {}), 12, True, True),
        "cv_soar_mp20": (("cv", (("single", (B.SOAR, 4000.0), {}), 3000.0), {}), 20, True, False),
        "multiple_mp16": (("multiple", (("single", (B.BARNES, 10000.0), {}), ("single", (B.BARNES, 100.0, 100.0), {}),
                                        ("single", (B.LINEAR, 0.5, 0.5, 0.5), {})), {}), 16, True, True),
        "barnes_hmax_mp40": (("single", (B.BARNES, 10000.0, 0.0, 0.0, 25000.0), {}), 40, True, False),
    }
    belev = rng.uniform(0, 500, (ny, nx)).astype(f32)
    blaf = rng.uniform(0, 1, (ny, nx)).astype(f32)
    pelev = rng.uniform(0, 500, S).astype(f32)
    plaf = rng.uniform(0, 1, S).astype(f32)
    pelev[11] = np.nan
    store = dict(y=yy.astype(f32), x=xx.astype(f32), background=bg, py=py, px=px, pobs=obs, pratios=ratios, pbackground=pbg,
                 belev=belev, blaf=blaf, pelev=pelev, plaf=plaf, names=np.array(sorted(cases)))
    for name, (spec, mp, extr, use_elev) in cases.items():
        s = build(np.array([repr(spec)]))
        be, bl, pe, pl = (belev, blaf, pelev, plaf) if use_elev else (None, None, None, None)
        out, var = ref.optimal_interpolation((yy, xx, be, bl), bg, (py, px, pe, pl), obs, ratios, pbg, s, mp, B.CARTESIAN,
                                             allow_extrapolation=extr, want_variance=True)
        store[name + "__structure"] = np.array([repr(spec)])
        store[name + "__args"] = np.array([mp, int(extr), int(use_elev)])
        store[name + "__analysis"] = out.reshape(ny, nx)
        store[name + "__variance"] = var.reshape(ny, nx)
    np.savez_compressed(os.path.join(HERE, "oi_c3_density.npz"), **store)

    # ---- EnSI: config 5 density on a 24x24 window, 6 members
    ny = nx = 24
    dx = 4000.0
    yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    S, E = 120, 6
    py, px = rng.uniform(0, ny * dx, S).astype(f32), rng.uniform(0, nx * dx, S).astype(f32)
    bgE = (rng.normal(size=(ny, nx, 1)) + rng.normal(size=(ny, nx, E))).astype(f32)
    pbgE = rng.normal(size=(S, E)).astype(f32)
    obsE = rng.normal(size=S).astype(f32)
    obsE[7] = np.nan
    sig = rng.uniform(0.3, 0.8, S).astype(f32)
    spec = ("single", (B.BARNES, 10000.0), {})
    s = build(np.array([repr(spec)]))
    store = dict(y=yy.astype(f32), x=xx.astype(f32), background=bgE, py=py, px=px, pobs=obsE, psigmas=sig, pbackground=pbgE,
                 structure=np.array([repr(spec)]))
    for name, mp, extr in (("mp20", 20, True), ("mp20_clamp", 20, False), ("unlimited", 0, True)):
        out = ref.optimal_interpolation_ensi((yy, xx, None, None), bgE, (py, px, None, None), obsE, sig, pbgE, s, mp, B.CARTESIAN,
                                             allow_extrapolation=extr)
        store[name + "__args"] = np.array([mp, int(extr)])
        store[name + "__analysis"] = out.reshape(ny, nx, E)
    np.savez_compressed(os.path.join(HERE, "ensi_c5_density.npz"), **store)

    # ---- neighbourhood / quantile_fast: config 2/4 style field with 1 % NaN, an all-NaN block and an inf
    f = (rng.uniform(size=(96, 130)) * 10).astype(f32)
    f[rng.uniform(size=f.shape) < 0.01] = np.nan
    f[40:62, 50:72] = np.nan
    f[5, 5] = np.inf
    store = dict(field=f)
    for hw in (0, 1, 7, 15, 70):
        for name, stat in (("mean", B.MEAN), ("sum", B.SUM), ("count", B.COUNT), ("min", B.MIN), ("max", B.MAX)):
            store["hw%d__%s" % (hw, name)] = ref.neighbourhood(f, hw, stat)
    np.savez_compressed(os.path.join(HERE, "neighbourhood.npz"), **store)

    thr = np.linspace(0, 10, 20).astype(f32)
    qfield = rng.uniform(size=f.shape).astype(f32)
    qfield[3, 3] = np.nan
    store = dict(field=f, thresholds=thr, qfield=qfield, thresholds_auto=ref.get_neighbourhood_thresholds(f, 11))
    for hw in (1, 7, 15):
        for q in (0.0, 0.001, 0.5, 0.9, 0.999, 1.0):
            store["hw%d__q%g" % (hw, q)] = ref.neighbourhood_quantile_fast(f, q, hw, thr)
    store["hw3__qfield"] = ref.neighbourhood_quantile_fast(f, qfield, 3, thr)
    store["hw2__auto_q0.9"] = ref.neighbourhood_quantile_fast(f, 0.9, 2, store["thresholds_auto"])
    np.savez_compressed(os.path.join(HERE, "quantile_fast.npz"), **store)

    # ---- index queries
    store = {}
    for tname, t in (("geodetic", B.GEODETIC), ("cartesian", B.CARTESIAN)):
        if t == B.CARTESIAN:
            la, lo = rng.uniform(0, 1e5, 1500).astype(f32), rng.uniform(0, 1e5, 1500).astype(f32)
            ql, qo = rng.uniform(-1e4, 1.1e5, 300).astype(f32), rng.uniform(-1e4, 1.1e5, 300).astype(f32)
            r = 8000.0
        else:
            la, lo = rng.uniform(55, 65, 1500).astype(f32), rng.uniform(0, 20, 1500).astype(f32)
            ql, qo = rng.uniform(54, 66, 300).astype(f32), rng.uniform(-1, 21, 300).astype(f32)
            r = 50000.0
        idx, dist, cnt = ref.points_neighbours(la, lo, t, ql, qo, r, capacity=96)
        x, y, z = ref.convert_coordinates(la, lo, t)
        vals = rng.normal(size=(3, 1500)).astype(f32)
        store.update({tname + "__lats": la, tname + "__lons": lo, tname + "__qlats": ql, tname + "__qlons": qo,
                      tname + "__radius": np.float32(r), tname + "__x": x, tname + "__y": y, tname + "__z": z,
                      tname + "__nearest": ref.points_nearest(la, lo, t, ql, qo),
                      tname + "__nearest_nomatch": ref.points_nearest(la, lo, t, la[:200], lo[:200], False),
                      tname + "__nbr_index": idx, tname + "__nbr_dist": dist, tname + "__nbr_count": cnt,
                      tname + "__closest5": ref.points_closest(la, lo, t, ql, qo, 5),
                      tname + "__values": vals, tname + "__nearest_values": ref.nearest(la, lo, t, ql, qo, vals)})
    np.savez_compressed(os.path.join(HERE, "index_queries.npz"), **store)

    # ---- structure functions
    n = 600
    p1 = np.zeros((n, 5), f32)
    p2 = np.zeros((n, 5), f32)
    p1[:, 0:2] = rng.uniform(0, 20000, (n, 2))
    p2[:, 0:2] = rng.uniform(0, 20000, (n, 2))
    p1[:, 3], p2[:, 3] = rng.uniform(0, 500, n), rng.uniform(0, 500, n)
    p1[:, 4], p2[:, 4] = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    p2[::7, 3] = np.nan
    p1[::11, 4] = np.nan
    store = dict(p1=p1, p2=p2)
    specs = {}
    for st, nm in enumerate(("barnes", "cressman", "soar", "toar", "powerlaw", "linear")):
        specs[nm] = ("single", (st, 5000.0, 200.0, 0.5), {})
        specs[nm + "_hmax"] = ("single", (st, 5000.0, 200.0, 0.5, 7000.0), {})
        specs[nm + "_cv"] = ("cv", (("single", (st, 5000.0, 200.0, 0.5), {}), 3000.0), {})
    specs["multiple"] = ("multiple", (("single", (B.BARNES, 5000.0), {}), ("single", (B.CRESSMAN, 300.0, 300.0, 300.0), {}),
                                      ("single", (B.LINEAR, 0.3, 0.3, 0.3), {})), {})
    store["names"] = np.array(sorted(specs))
    for nm, spec in specs.items():
        s = build(np.array([repr(spec)]))
        store[nm + "__structure"] = np.array([repr(spec)])
        store[nm + "__corr"] = ref.structure_corr(s, p1, p2, False)
        store[nm + "__corr_background"] = ref.structure_corr(s, p1, p2, True)
        store[nm + "__loc_dist"] = np.float32(ref.structure_localization_distance(s))
    np.savez_compressed(os.path.join(HERE, "structure.npz"), **store)
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
