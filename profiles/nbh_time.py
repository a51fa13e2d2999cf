"""Times the neighbourhood kernels device-resident (CUDA events, inputs rotating over 4 buffers = 256 MB > L2).
usage: python profiles/nbh_time.py [n] ; GPP_NO_TMA=1 selects the plain kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
bufs = [torch.rand((n, n), device="cuda") * 10 for _ in range(4)]
out = torch.empty((n, n), device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=20):
    for i in range(3):
        fn(bufs[i % 4])
    ev0.record()
    for i in range(reps):
        fn(bufs[i % 4])
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


ms = timeit(lambda b: out.copy_(b))
print("torch copy_ (same traffic, the practical ceiling at this size)  %.4f ms  %.0f GB/s" % (ms, 8.0 * n * n / ms / 1e6))
for hw in (7, 15):
    for name, st in (("mean", gpp.Mean), ("count", gpp.Count), ("min", gpp.Min), ("max", gpp.Max)):
        ms = timeit(lambda b: gd.neighbourhood(b, hw, st, out=out))
        print("%-6s hw=%2d n=%d  %.4f ms  %.0f GB/s" % (name, hw, n, ms, 8.0 * n * n / ms / 1e6))
nanbuf = bufs[0].clone()
nanbuf[torch.rand((n, n), device="cuda") < 0.01] = float("nan")
ms = timeit(lambda b: gd.neighbourhood(nanbuf, 7, gpp.Mean, out=out))
print("mean hw=7 with 1%% NaN (L2-resident input)  %.4f ms  %.0f GB/s" % (ms, 8.0 * n * n / ms / 1e6))
if "--qf" in sys.argv:
    thr = np.linspace(0, 10, 20).astype(np.float32)
    for hw, T in ((7, 11), (15, 20)):
        t = np.linspace(0, 10, T).astype(np.float32)
        ms = timeit(lambda b: gd.neighbourhood_quantile_fast(b, 0.5, hw, t, out=out), reps=4)
        print("qfast  hw=%2d T=%d n=%d  %.4f ms  %.0f GB/s" % (hw, T, n, ms, 8.0 * n * n / ms / 1e6))
