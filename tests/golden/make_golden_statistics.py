"""Golden fixture for the statistics family of round 2, generated from the UNMODIFIED reference sources
(oracle/_ref/libgridpp_ref.so), like make_golden.py:

    make -C oracle ref && python tests/golden/make_golden_statistics.py

gridpp::neighbourhood with Std / Variance / Median (neighbourhood.cpp:211-238), gridpp::neighbourhood_brute_force and
gridpp::neighbourhood_quantile for vec2 and vec3 (:528-630), gridpp::calc_statistic / calc_quantile (util.cpp:19-215)
and gridpp::interpolate (util.cpp:377-431).
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
    ny, nx, ne = 48, 61, 5
    f = (rng.gamma(0.5, 2.0, size=(ny, nx)) * 3).astype(f32)
    f[rng.uniform(size=f.shape) < 0.05] = np.nan
    f[10:16, 20:27] = np.nan                               # windows without a valid value
    f[rng.uniform(size=f.shape) < 0.2] = 0                 # ties
    e = (f[:, :, None] + rng.normal(size=(ny, nx, ne))).astype(f32)
    e[rng.uniform(size=e.shape) < 0.03] = np.nan
    store = dict(field=f, ensemble=e)
    for hw in (0, 1, 3, 7):
        for name in ("std", "variance", "median"):
            store["nbh_hw%d__%s" % (hw, name)] = ref.neighbourhood(f, hw, STATS[name])
    for hw in (1, 4):
        for name, st in STATS.items():
            store["brute_hw%d__%s" % (hw, name)] = ref.neighbourhood_window(f, hw, st)
            store["brute_ens_hw%d__%s" % (hw, name)] = ref.neighbourhood_window(e, hw, st)
    for hw in (0, 2, 5):
        for q in (0.0, 0.1, 0.5, 0.77, 1.0):
            store["quantile_hw%d__q%g" % (hw, q)] = ref.neighbourhood_window(f, hw, B.QUANTILE, q)
            store["quantile_ens_hw%d__q%g" % (hw, q)] = ref.neighbourhood_window(e, hw, B.QUANTILE, q)
    # rows for calc_statistic / calc_quantile
    rows = (rng.normal(size=(300, 23)) * 5 + 100).astype(f32)
    rows[rng.uniform(size=rows.shape) < 0.1] = np.nan
    rows[5] = np.nan
    rows[6, 1:] = np.nan
    rows[7] = 3.25
    store["rows"] = rows
    for name, st in STATS.items():
        store["rows__" + name] = np.array([ref.calc_statistic(r, st) for r in rows], f32)
    for q in (0.0, 0.25, 0.5, 0.9, 1.0):
        store["rows__q%g" % q] = np.array([ref.calc_quantile(r, q) for r in rows], f32)
    ix = np.array([0, 0, 0.5, 0.5, 1, 1, 2.5, 4], f32)
    iy = np.array([0, 0.1, 0.4, 0.6, 0.9, 1, 3, 2], f32)
    x = np.concatenate([rng.uniform(-1, 5, 200), ix, [np.nan]]).astype(f32)
    store.update(interp_ix=ix, interp_iy=iy, interp_x=x, interp_y=np.array([ref.interpolate(v, ix, iy) for v in x], f32))
    np.savez_compressed(os.path.join(HERE, "statistics.npz"), **store)
    print("wrote statistics.npz (%d arrays)" % len(store))


if __name__ == "__main__":
    main()
