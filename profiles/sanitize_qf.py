"""Tiny quantile_fast / min-max invocations for compute-sanitizer --tool racecheck (the full sanitize_all.py takes ~14
minutes under racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp

rng = np.random.default_rng(0)
f32 = np.float32
for shape in ((20, 260), (18, 77)):
    f = rng.uniform(size=shape).astype(f32)
    f[rng.uniform(size=shape) < 0.03] = np.nan
    f[3, 5] = np.inf
    f[4, shape[1] - 1] = -np.inf
    thr = np.linspace(0, 1, 20).astype(f32)
    for hw in (2, 15):
        gpp.neighbourhood_quantile_fast(f, 0.5, hw, thr)
        gpp.neighbourhood_quantile_fast(f, rng.uniform(size=shape).astype(f32), hw, thr[:7])
        gpp.neighbourhood(f, hw, gpp.Min)
        gpp.neighbourhood(f, hw, gpp.Mean)
gpp.synchronize()
print("sanitize_qf: done")
