#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GPP_TRACE=1 timeout 600 python -c "
import json, time, numpy as np, bench, gridpp_b200 as gpp
w = bench.ensi_inputs(0, bench.ENSI_N)
n, E = bench.ENSI_N, bench.ENSI_E
grid, points = gpp.Grid(w['y'], w['x'], type=gpp.Cartesian), gpp.Points(w['py'], w['px'], type=gpp.Cartesian)
s = gpp.BarnesStructure(bench.H_SCALE)
pobs = (w['obs'][:, None] + 0.1 * np.arange(E, dtype=np.float32)[None, :]).astype(np.float32)
pr = np.full(bench.ENSI_S, 0.25, np.float32); br = np.ones((n, n), np.float32)
for i in range(2):
    t0 = time.perf_counter()
    out = gpp.optimal_interpolation_ensi_multi_ebesc(grid, br, w['bg'], points, pobs, pr, w['pbg'], s, 30, False)
    print('ebesc call', time.perf_counter() - t0)
for i in range(2):
    t0 = time.perf_counter()
    out = gpp.optimal_interpolation_ensi_multi_ebe(grid, br, w['bg'], w['bg'], points, pobs, pr, w['pbg'], w['pbg'], s, 30, False)
    print('ebe call', time.perf_counter() - t0)
" 2>&1 | tail -24
