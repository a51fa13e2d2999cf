"""Launches each neighbourhood kernel a few times at config-2 size (4000 x 4000, halfwidth 7) for ncu.
usage: python profiles/nbh_probe.py [mean|min|qf ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

which = sys.argv[1:] or ["mean", "min"]
n = 4000
bufs = [torch.rand((n, n), device="cuda") * 10 for _ in range(3)]
out = torch.empty((n, n), device="cuda")
thr = np.linspace(0, 10, 20).astype(np.float32)
for i in range(3):
    if "mean" in which:
        gd.neighbourhood(bufs[i], 7, gpp.Mean, out=out)
    if "min" in which:
        gd.neighbourhood(bufs[i], 7, gpp.Min, out=out)
    if "qf" in which:
        gd.neighbourhood_quantile_fast(bufs[i], 0.5, 15, thr, out=out)
torch.cuda.synchronize()
