"""Times the OI kernels beyond the register path (max_points > 30: shared-memory Cholesky up to 128 observations per
point; the general global-memory kernel beyond that or for non-symmetric structure functions) on the C3 workload.
usage: python profiles/oi_general_time.py [rows]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = bench.make_workload(0, rows)
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
bg = torch.from_numpy(w["background"].ravel()).cuda()
out = torch.empty_like(bg)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, s, mp in (("Barnes mp=30 (register kernel)", gpp.BarnesStructure(bench.H_SCALE), 30),
                    ("Barnes mp=50 (Cholesky kernel, k <= 64)", gpp.BarnesStructure(bench.H_SCALE), 50),
                    ("Barnes unlimited (Cholesky kernel, k <= 128)", gpp.BarnesStructure(bench.H_SCALE), 0),
                    ("Cressman mp=30 (register kernel: no elevations)", gpp.CressmanStructure(36000.0), 30),
                    ("Soar h=10 km, v=200 m with elevations, mp=30 (general kernel: P not symmetric)", "soar", 30)):
    if s == "soar":
        rng = np.random.default_rng(3)
        grid = gpp.Grid(w["y"], w["x"], rng.uniform(0, 300, w["y"].shape).astype(np.float32), type=gpp.Cartesian)
        points = gpp.Points(w["py"], w["px"], rng.uniform(0, 300, w["py"].size).astype(np.float32), type=gpp.Cartesian)
        s = gpp.SoarStructure(bench.H_SCALE, 200.0)
    state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
    gd.optimal_interpolation(grid, bg, state, mp, out=out)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        gd.optimal_interpolation(grid, bg, state, mp, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    print("%-55s %8.2f ms  %7.1f M gridpoints/s" % (name, ms, rows * bench.N_GRID / ms / 1e3))
