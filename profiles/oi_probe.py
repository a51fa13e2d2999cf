"""Launches the C3 OI analysis a few times (device-resident) for ncu. usage: python profiles/oi_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

w = bench.make_workload()
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s = gpp.BarnesStructure(bench.H_SCALE)
state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
bg = torch.from_numpy(w["background"].ravel()).cuda()
out = torch.empty_like(bg)
for _ in range(3):
    gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
torch.cuda.synchronize()
