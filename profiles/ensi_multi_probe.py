"""Launches optimal_interpolation_ensi_multi_ebesc on 400 rows of config 5's geometry (1 M points, 20 members, max_points 30)
for ncu: python profiles/ensi_multi_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp

rows = 400
w = bench.ensi_inputs(0, rows)
E = bench.ENSI_E
grid, points = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian), gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
pobs = (w["obs"][:, None] + 0.1 * np.arange(E, dtype=np.float32)[None, :]).astype(np.float32)
for _ in range(2):
    out = gpp.optimal_interpolation_ensi_multi_ebesc(grid, np.ones((rows, bench.ENSI_N), np.float32), w["bg"], points, pobs,
                                                     np.full(bench.ENSI_S, 0.25, np.float32), w["pbg"], gpp.BarnesStructure(bench.H_SCALE), 30, False)
print("done", float(np.nansum(out)))
