import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench, gridpp_b200 as gpp
w = bench.make_workload()
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian); points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian); s = gpp.BarnesStructure(10000.)
h_bg = torch.from_numpy(w["background"]).pin_memory().numpy()
for i in range(3):
    t0=time.perf_counter(); out = gpp.optimal_interpolation(grid, h_bg, points, w["pobs"], w["pratios"], w["pbackground"], s, 30); print("call ms", 1e3*(time.perf_counter()-t0), file=sys.stderr)
