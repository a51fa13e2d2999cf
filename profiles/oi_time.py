"""Device-resident time of the C3 OI analysis (CUDA events), plus the reuse statistics when the library was built with
-DOI_STATS. usage: [GPP_B200_LIB=scratch/lib_x.so] python profiles/oi_time.py [rows]
rows < 4000 analyses only the first `rows` rows of the 4000 x 4000 grid: rows = 500 is what one rank of an 8-GPU run does,
so (time of 4000 rows) / 8 / (time of 500 rows) is the strong-scaling efficiency of the kernel itself."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import _lib, device as gd

rows = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_GRID
w = bench.make_workload(0, rows)
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s = gpp.BarnesStructure(bench.H_SCALE)
state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
bg = torch.from_numpy(w["background"].ravel()).cuda()
out = torch.empty_like(bg)
for _ in range(3):
    gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record()
    gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
ev[5].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
line = "%s: %d rows: %.3f ms (min %.3f) = %.1f M gridpoints/s, checksum %.6f" % (os.environ.get("GPP_B200_LIB", "default"), rows, sum(ms) / 5, min(ms),
                                                                             rows * bench.N_GRID / min(ms) / 1e3, float(torch.nan_to_num(out).double().sum()))
fn = getattr(_lib.lib, "gpp_debug_oi_stats", None) if hasattr(_lib.lib, "gpp_debug_oi_stats") else None
if fn is not None:
    stats = (ctypes.c_ulonglong * 4)()
    fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    fn(stats, 1)
    gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
    fn(stats, 1)
    line += "; selection changes %.1f%%, systems solved %.2f%%, full rankings %.2f%% of points" % (
        100.0 * stats[1] / stats[0], 100.0 * stats[2] / stats[0], 100.0 * stats[3] / stats[0])
print(line)
