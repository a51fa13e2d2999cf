"""Golden fixture for the ensemble (vec3) forms of the neighbourhood filters, generated from the UNMODIFIED reference
sources (oracle/_ref/libgridpp_ref.so), like make_golden.py:

    make -C oracle ref && python tests/golden/make_golden_ensemble.py

gridpp::neighbourhood(vec3, ...) neighbourhood.cpp:12-27, gridpp::neighbourhood_quantile_fast(vec3, ...) :411-527.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)
    rng = np.random.default_rng(1000)
    f32 = np.float32
    ny, nx, ne = 60, 72, 7
    f = (rng.gamma(0.5, 2.0, size=(ny, nx, 1)) + rng.normal(size=(ny, nx, ne)) * 0.5).astype(f32)
    f[rng.uniform(size=f.shape) < 0.05] = np.nan          # single missing members
    f[20:30, 30:45, :] = np.nan                           # cells without any valid member
    f[rng.uniform(size=f.shape) < 0.2] = 0                # exact zeros: plateaus
    store = dict(field=f)
    for hw in (0, 1, 4, 9):
        for name, stat in (("mean", B.MEAN), ("sum", B.SUM), ("count", B.COUNT), ("min", B.MIN), ("max", B.MAX)):
            store["nbh_hw%d__%s" % (hw, name)] = ref.neighbourhood_ens(f, hw, stat)
    thr = np.array([0, 0.1, 0.5, 1, 2, 4, 8], f32)
    qfield = rng.uniform(size=(ny, nx)).astype(f32)
    qfield[2, 2] = np.nan
    store.update(thresholds=thr, qfield=qfield)
    for hw in (0, 2, 6):
        for q in (0.0, 0.3, 0.5, 0.9, 1.0):
            store["qf_hw%d__q%g" % (hw, q)] = ref.neighbourhood_quantile_fast_ens(f, q, hw, thr)
    store["qf_hw3__qfield"] = ref.neighbourhood_quantile_fast_ens(f, qfield, 3, thr)
    np.savez_compressed(os.path.join(HERE, "ensemble_forms.npz"), **store)
    print("wrote ensemble_forms.npz (%d arrays)" % len(store))


if __name__ == "__main__":
    main()
