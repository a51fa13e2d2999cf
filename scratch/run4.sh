#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest4.log
tail -12 $O/r2_pytest4.log | cut -c1-300
{
python profiles/nbh_time.py 4000 --qf
python profiles/nbh_time.py 8000
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, bench
import gridpp_b200 as gpp
from gridpp_b200 import device as gd
w = bench.make_workload(0, 1000)
grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian); points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
s = gpp.BarnesStructure(bench.H_SCALE)
state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], s)
bg = torch.from_numpy(w["background"].ravel()).cuda(); out = torch.empty_like(bg); var = torch.empty_like(bg)
for name, v in (("analysis only", None), ("analysis + variance", var)):
    gd.optimal_interpolation(grid, bg, state, 30, out=out, out_variance=v); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): gd.optimal_interpolation(grid, bg, state, 30, out=out, out_variance=v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("OI 1000 rows, %s: %.2f ms = %.1f M gridpoints/s" % (name, ms, 4e6 / ms / 1e3))
PY
} > $O/r2_time4.log 2>&1
cat $O/r2_time4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbh_sumf_tma_kernel -s 2 -c 1 -o $O/r2_nbh_mean_v3 -f python profiles/nbh_probe.py mean > $O/r2_ncu4.log 2>&1
tail -2 $O/r2_ncu4.log
