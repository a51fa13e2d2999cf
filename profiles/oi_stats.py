"""Reuse statistics of the C3 OI analysis (needs a library built with -DOI_STATS: profiles/variants.sh stats "-DOI_STATS";
GPP_B200_LIB=$PWD/scratch/lib_stats.so python profiles/oi_stats.py)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import _lib, device as gd

w = bench.make_workload()
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s = gpp.BarnesStructure(bench.H_SCALE)
state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
bg = torch.from_numpy(w["background"].ravel()).cuda()
out = torch.empty_like(bg)
stats = (ctypes.c_ulonglong * 4)()
fn = _lib.lib.gpp_debug_oi_stats
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
fn(stats, 1)
gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
fn(stats, 1)
pts, changes, solves = stats[0], stats[1], stats[2]
print("points on the run path %d, selection changes %d (%.1f%%), systems solved %d (%.1f%% of points, %.1f%% of changes)"
      % (pts, changes, 100.0 * changes / pts, solves, 100.0 * solves / pts, 100.0 * solves / max(changes, 1)))
