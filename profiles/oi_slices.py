"""Strong-scaling emulation on ONE GPU: the C3 grid cut into N row blocks exactly as bench.py --gpus N cuts it, each block timed
alone (CUDA events, L2 flushed between launches as bench.py does for blocks smaller than L2). max over blocks = the step time an
N-GPU run reports; whole / N / max = its efficiency. usage: python profiles/oi_slices.py [N ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

parts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
s = gpp.BarnesStructure(bench.H_SCALE)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
whole = None
for n in parts:
    times = []
    for r in range(n):
        row0, row1 = bench.N_GRID * r // n, bench.N_GRID * (r + 1) // n
        w = bench.make_workload(row0, row1)
        grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
        points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
        state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
        bg = torch.from_numpy(w["background"].ravel()).cuda()
        out = torch.empty_like(bg)
        for _ in range(3):
            gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
        ms = []
        for i in range(5):
            flush.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gd.optimal_interpolation(grid, bg, state, bench.MAX_POINTS, out=out)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        times.append(sum(ms) / len(ms))
    if n == 1:
        whole = times[0]
    line = "N=%d: blocks (ms) %s  max %.3f" % (n, " ".join("%.3f" % t for t in times), max(times))
    if whole:
        line += "  efficiency %.3f (speed-up %.2f)" % (whole / n / max(times), whole / max(times))
    print(line, flush=True)
