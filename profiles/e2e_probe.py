"""End-to-end time of the C3 OI analysis through the host API (pinned input, fresh pageable output), with the phase
trace of the library. usage: [GPP_OI_CHUNKS=n] [GPP_TRACE=1] python profiles/e2e_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp

rows = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_GRID     # rows < 4000: what one rank of a multi-GPU run analyses
w = bench.make_workload(0, rows)
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s = gpp.BarnesStructure(10000.)
h_bg = torch.from_numpy(w["background"]).pin_memory().numpy()
times = []
for i in range(8):
    t0 = time.perf_counter()
    out = gpp.optimal_interpolation(grid, h_bg, points, w["pobs"], w["pratios"], w["pbackground"], s, 30)
    times.append(1e3 * (time.perf_counter() - t0))
print("rows %d, chunks %s, pipeline_min %s: call ms %s; best %.2f" % (rows, os.environ.get("GPP_OI_CHUNKS", "default"), os.environ.get("GPP_OI_PIPELINE_MIN", "default"), " ".join("%.1f" % t for t in times), min(times)))
