"""Device-resident time of the C5 EnSI analysis (gpp_optimal_interpolation_ensi_device, CUDA events), for variant
builds: [GPP_B200_LIB=scratch/lib_x.so] python profiles/ensi_device_time.py [rows] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import device as gd

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, dx, E, S = 2500, 200.0, (int(sys.argv[3]) if len(sys.argv) > 3 else 20), 5000
rng = np.random.default_rng(1000)
y, x = np.meshgrid(np.arange(rows, dtype=np.float32) * dx, np.arange(n, dtype=np.float32) * dx, indexing="ij")
py, px = (rng.random(S) * n * dx).astype(np.float32), (rng.random(S) * n * dx).astype(np.float32)
pbg = rng.standard_normal((S, E)).astype(np.float32)
obs = rng.standard_normal(S).astype(np.float32)
sig = np.full(S, 0.5, np.float32)
grid, points = gpp.Grid(y, x, type=gpp.Cartesian), gpp.Points(py, px, type=gpp.Cartesian)
s = gpp.BarnesStructure(10000)
g = torch.Generator(device="cuda").manual_seed(1000)
bg = torch.randn((rows, n, E), device="cuda", generator=g) + 2 * torch.randn((rows, n, 1), device="cuda", generator=g)
state = gd.EnsembleObservationState(points, obs, sig, pbg, s)
out = torch.empty_like(bg)
gd.optimal_interpolation_ensi(grid, bg, state, 50, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    gd.optimal_interpolation_ensi(grid, bg, state, 50, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print("%s: EnSI %d x %d x %d, mp 50: %.1f ms (min %.1f) = %.2f M gridpoints/s, checksum %.6f" % (
    os.environ.get("GPP_B200_LIB", "default"), rows, n, E, sum(ms) / reps, min(ms), rows * n / min(ms) / 1e3,
    float(out.double().sum())))
ref = "/tmp/ensi_device_time_ref_%d_%d.pt" % (rows, E)
if os.path.exists(ref):
    want = torch.load(ref)
    err = ((out.cpu() - want).abs() / want.abs().clamp(min=1.0)).max().item()
    print("   max |diff| / max(|default|, 1) against the first run of this session: %.3e" % err)
else:
    torch.save(out.cpu(), ref)
