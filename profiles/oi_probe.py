"""Launches the C3 OI analysis a few times (device-resident) for ncu.
usage: python profiles/oi_probe.py [fast | chol | general | full]   (chol: max_points 50 on 1000 rows; general: Soar with
elevations on 250 rows; full: the analysis-variance path of the register kernel on 1000 rows)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

import numpy as np

which = sys.argv[1] if len(sys.argv) > 1 else "fast"
rows = {"fast": bench.N_GRID, "chol": 1000, "general": 250, "full": 1000}[which]
w = bench.make_workload(0, rows)
mp = 50 if which == "chol" else bench.MAX_POINTS
if which == "general":   # Soar with an active vertical scale is not symmetric in the reference -> general kernel
    rng = np.random.default_rng(3)
    grid = gpp.Grid(w["y"], w["x"], rng.uniform(0, 300, w["y"].shape).astype(np.float32), type=gpp.Cartesian)
    points = gpp.Points(w["py"], w["px"], rng.uniform(0, 300, w["py"].size).astype(np.float32), type=gpp.Cartesian)
    s = gpp.SoarStructure(bench.H_SCALE, 200.0)
else:
    grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
    points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
    s = gpp.BarnesStructure(bench.H_SCALE)
state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
bg = torch.from_numpy(w["background"].ravel()).cuda()
out = torch.empty_like(bg)
var = torch.empty_like(bg) if which == "full" else None
for _ in range(3):
    gd.optimal_interpolation(grid, bg, state, mp, out=out, out_variance=var)
torch.cuda.synchronize()
