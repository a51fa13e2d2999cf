"""Golden fixture for the last two window filters of SURVEY.md 8(f)#4, generated from the UNMODIFIED reference sources
(oracle/_ref/libgridpp_ref.so):

    make -C oracle ref && python tests/golden/make_golden_window_filters.py

gridpp::neighbourhood_search (neighbourhood_search.cpp) and gridpp::calc_gradient (calc_gradient.cpp, MinMax and LinearRegression).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
SEARCH_CASES = [(1, 0.5, 1.0, 0.1), (3, 0.8, 1.0, 0.0), (2, 0.0, 0.05, 0.3), (4, 2.0, 3.0, 0.2), (0, 0.2, 0.9, 0.0)]
GRADIENT_CASES = [(1, 0, np.nan, 0.0), (2, 2, np.nan, -11.0), (2, 5, 0.0, -11.0), (4, 2, 0.5, 3.0), (7, 30, 0.2, 0.0)]


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)
    rng = np.random.default_rng(77)
    ny, nx = 40, 52
    temp = (10 + 3 * rng.standard_normal((ny, nx))).astype(f32)
    temp[rng.uniform(size=temp.shape) < 0.04] = np.nan
    laf = rng.uniform(0, 1, (ny, nx)).astype(f32)
    laf[rng.uniform(size=laf.shape) < 0.3] = 1.0
    laf[rng.uniform(size=laf.shape) < 0.03] = np.nan
    laf[10:20, 5:30] = 0.0                                 # a lake: no cell of the target range in reach for small windows
    apply = (rng.uniform(size=(ny, nx)) < 0.7).astype(np.int32)
    apply[3, :] = 2                                        # neither 0 nor 1: treated, but no neighbour is ever used (:63)
    store = dict(search__array=temp, search__search=laf, search__apply=apply, search__cases=np.array(SEARCH_CASES, np.float64),
                 gradient__cases=np.array(GRADIENT_CASES, np.float64))
    for i, (hw, lo, hi, delta) in enumerate(SEARCH_CASES):
        store["search__case%d" % i] = ref.neighbourhood_search(temp, laf, hw, lo, hi, delta)
        store["search__case%d_apply" % i] = ref.neighbourhood_search(temp, laf, hw, lo, hi, delta, apply)
    elev = (300 + 200 * rng.standard_normal((ny, nx))).astype(f32)
    elev[rng.uniform(size=elev.shape) < 0.05] = np.nan
    elev[25:30, 25:30] = 150.0                             # a plateau: no range of the base
    t2m = (15 - 0.0065 * np.nan_to_num(elev, nan=300.0) + 0.5 * rng.standard_normal((ny, nx))).astype(f32)
    t2m[rng.uniform(size=t2m.shape) < 0.05] = np.nan
    store.update(gradient__base=elev, gradient__values=t2m)
    for i, (hw, num_min, min_range, default) in enumerate(GRADIENT_CASES):
        for name, gt in (("minmax", 0), ("regression", 10)):
            store["gradient__%s_case%d" % (name, i)] = ref.calc_gradient(elev, t2m, gt, hw, num_min, min_range, default)
    np.savez_compressed(os.path.join(HERE, "window_filters.npz"), **store)
    print("wrote window_filters.npz (%d arrays)" % len(store))


if __name__ == "__main__":
    main()
