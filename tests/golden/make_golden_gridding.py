"""Golden fixture for the consumers of the point index, generated from the UNMODIFIED reference sources
(oracle/_ref/libgridpp_ref.so), like make_golden.py:

    make -C oracle ref && python tests/golden/make_golden_gridding.py

gridpp::gridding / gridding_nearest (gridding.cpp), count (count.cpp), distance (distance.cpp), fill / fill_missing (fill.cpp),
doping_square / doping_circle (doping.cpp), for a Cartesian and a Geodetic set-up.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STATS = dict(mean=B.MEAN, min=B.MIN, median=B.MEDIAN, max=B.MAX, std=B.STD, variance=B.VARIANCE, sum=B.SUM, count=B.COUNT)


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)
    rng = np.random.default_rng(1000)
    f32 = np.float32
    store = {}
    for tag, ctype in (("cart", B.CARTESIAN), ("geo", B.GEODETIC)):
        ny, nx, S = 36, 44, 150
        if ctype == B.CARTESIAN:
            gy, gx = np.meshgrid(np.arange(ny) * 1000.0, np.arange(nx) * 1000.0, indexing="ij")
            pl, po = rng.uniform(-2000, ny * 1000 + 2000, S), rng.uniform(-2000, nx * 1000 + 2000, S)
            radius = 3500.0
        else:
            gy, gx = np.meshgrid(np.linspace(59, 60.5, ny), np.linspace(9, 12, nx), indexing="ij")
            pl, po = rng.uniform(58.9, 60.6, S), rng.uniform(8.9, 12.1, S)
            radius = 12000.0
        gy, gx, pl, po = gy.astype(f32), gx.astype(f32), pl.astype(f32), po.astype(f32)
        pl[10], po[10] = pl[11], po[11]                                  # a duplicate location
        values = rng.normal(size=S).astype(f32) * 3
        values[rng.uniform(size=S) < 0.05] = np.nan
        ge = rng.uniform(0, 600, (ny, nx)).astype(f32)
        pe = rng.uniform(0, 600, S).astype(f32)
        field = rng.normal(size=(ny, nx)).astype(f32)
        radii = rng.uniform(0, radius, S).astype(f32)
        radii[:5] = 0
        hw = rng.integers(0, 4, S).astype(np.int32)
        opl, opo = gy.ravel()[::7] + f32(0.01 if ctype == B.GEODETIC else 130.0), gx.ravel()[::7].copy()   # an output Points set
        store.update({tag + "__" + k: v for k, v in dict(glats=gy, glons=gx, gelevs=ge, plats=pl, plons=po, pelevs=pe, values=values, field=field,
                                                        radii=radii, halfwidth=hw, olats=opl, olons=opo, radius=f32(radius)).items()})
        for name, st in STATS.items():
            for mn in (0, 3):
                store["%s__gridding_grid__%s_mn%d" % (tag, name, mn)] = ref.gridding((gy, gx), (pl, po), values, radius, mn, st, ctype)
                store["%s__gridding_nearest_grid__%s_mn%d" % (tag, name, mn)] = ref.gridding((gy, gx), (pl, po), values, 0, mn, st, ctype, nearest=True)
            store["%s__gridding_points__%s" % (tag, name)] = ref.gridding((opl, opo), (pl, po), values, radius, 1, st, ctype)
            store["%s__gridding_nearest_points__%s" % (tag, name)] = ref.gridding((opl, opo), (pl, po), values, 0, 0, st, ctype, nearest=True)
        sets = dict(grid=(gy, gx), points=(pl, po), ogrid=(gy[:9, :11] + f32(0.004 if ctype == B.GEODETIC else 400.0), gx[:9, :11]), opoints=(opl, opo))
        for i, o in (("grid", "points"), ("grid", "ogrid"), ("points", "grid"), ("points", "opoints")):
            store["%s__count__%s_%s" % (tag, i, o)] = ref.count(sets[i], sets[o], radius, ctype)
            for num in (1, 4):
                store["%s__distance__%s_%s_n%d" % (tag, i, o, num)] = ref.distance(sets[i], sets[o], num, ctype)
        for outside in (0, 1):
            store["%s__fill__outside%d" % (tag, outside)] = ref.fill(gy, gx, field, pl, po, radii, -7.5, outside, ctype)
        for med, mname in ((np.nan, "nocheck"), (150.0, "elev150")):
            store["%s__doping_circle__%s" % (tag, mname)] = ref.doping(gy, gx, ge, field, pl, po, pe, values, radii, med, ctype, False)
            store["%s__doping_square__%s" % (tag, mname)] = ref.doping(gy, gx, ge, field, pl, po, pe, values, hw, med, ctype, True)
    fm = rng.normal(size=(40, 53)).astype(f32)
    fm[rng.uniform(size=fm.shape) < 0.3] = np.nan
    fm[5, :] = np.nan
    fm[:, 7] = np.nan
    fm[20:30, 10:25] = np.nan
    store["fill_missing__in"] = fm
    store["fill_missing__out"] = ref.fill_missing(fm)
    np.savez_compressed(os.path.join(HERE, "gridding.npz"), **store)
    print("wrote gridding.npz (%d arrays)" % len(store))


if __name__ == "__main__":
    main()
