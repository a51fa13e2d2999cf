#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native gridpp hot path.

Metric (BASELINE.json): OI analysis gridpoints/sec on config 3 -- optimal_interpolation over a 4000 x 4000
Cartesian grid (dx 250 m), 10 000 uniformly placed observations, BarnesStructure(10 km), variance ratio 0.5,
max_points = 30.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]     the reference's own CPU implementation

One "step" = one full analysis of the grid. For N > 1 (launched by torchrun, one rank per GPU) the output rows
are split across ranks (every grid point is independent, oi.cpp:221-338): no data-path collective, the
observations are replicated; `value` = total gridpoints / max-over-ranks device time (strong scaling of the
fixed 4000 x 4000 grid that the metric names).

Printed JSON (one line, rank 0):
  value      device-resident throughput: background / analysis already in HBM, CUDA-event timed.
  e2e        the same analysis through the public host API (gridpp_b200.optimal_interpolation -> C ABI
             gpp_optimal_interpolation_host) from pinned HOST arrays: H2D of the background, kernels, D2H of
             the analysis are all inside the timed region.
  roofline   dominant kernel (oi_fast_kernel). The kernel is bound by the fp64 CUDA cores, not by HBM or the
             tensor cores; the contract's object reports the HBM view (16 B/gridpoint algorithmic traffic),
             `roofline_fp64` reports the algorithmic-flop view against a measured fp64 FMA peak.
  cpu_baseline  the reference's CPU path (oracle/_ref: the unmodified reference sources, all host threads) on
             a bounded row-strided sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner at most
# NCCL_DEBUG levels). Keep a private handle on the real stdout for the result and point fd 1 at stderr.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


N_GRID = 4000
DX = 250.0
N_OBS = 10000
H_SCALE = 10000.0
MAX_POINTS = 30
RATIO = 0.5
SEED = 1000
WORKLOAD = "C3: optimal_interpolation 4000x4000 grid (dx 250 m), 10000 obs, BarnesStructure(10000), ratio 0.5, max_points 30"
BACKGROUND = ("analytic 3 sin(y/37 km) cos(x/53 km) + 1.5 sin((y+x)/11 km) instead of SURVEY 8d's gaussian-filtered noise: every rank "
              "can evaluate its own rows and the values at the observation points without the whole field; the cost of OI does not "
              "depend on the background values")


def make_workload(row0=0, row1=N_GRID):
    """Deterministic synthetic inputs; rows [row0, row1) of the grid. Observations are global."""
    rng = np.random.default_rng(SEED)
    py = (rng.random(N_OBS) * N_GRID * DX).astype(np.float32)
    px = (rng.random(N_OBS) * N_GRID * DX).astype(np.float32)
    noise = rng.standard_normal(N_OBS).astype(np.float32) * 0.5
    ys = np.arange(row0, row1, dtype=np.float32) * DX
    xs = np.arange(N_GRID, dtype=np.float32) * DX
    y, x = np.meshgrid(ys, xs, indexing="ij")
    # smooth synthetic background (analytic, so every rank can evaluate its own rows and the obs-point values)
    def field(yy, xx):
        return (3.0 * np.sin(yy / 37000.0) * np.cos(xx / 53000.0) + 1.5 * np.sin((yy + xx) / 11000.0)).astype(np.float32)
    bg = field(y, x)
    # background at the observation points = value at the nearest grid node (what gridpp.nearest returns)
    iy = np.clip(np.rint(py / DX), 0, N_GRID - 1).astype(np.float32) * DX
    ix = np.clip(np.rint(px / DX), 0, N_GRID - 1).astype(np.float32) * DX
    pbg = field(iy, ix)
    obs = (pbg + noise).astype(np.float32)
    ratios = np.full(N_OBS, RATIO, np.float32)
    return dict(y=y, x=x, background=bg, py=py, px=px, pobs=obs, pratios=ratios, pbackground=pbg)


def oi_flops_per_gridpoint(w):
    """F_OI(N_c, k) = 20 (N_c + k(k+1)/2) + 2 (k^3/3 + 2 k^2) + 4 k, averaged over a sample."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(1)
    R = float(np.sqrt(np.float32(-2) * np.log(np.float32(0.0013)))) * H_SCALE
    tree = cKDTree(np.stack([w["py"], w["px"]], axis=1))
    q = rng.random((20000, 2)) * N_GRID * DX
    nc = np.array([len(v) for v in tree.query_ball_point(q, R)], dtype=np.float64)
    k = np.minimum(nc, MAX_POINTS)
    f = 20 * (nc + k * (k + 1) / 2) + 2 * (k ** 3 / 3 + 2 * k ** 2) + 4 * k
    return float(f.mean()), float(nc.mean()), float(k.mean())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (guides/B200_PROFILING.md)."""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.samples, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([p.strip() for p in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_run(w, n_sample, threads=None):
    """Times the reference's CPU path (oracle/_ref when present, else the C port) on a row-strided sample of the grid
    (Points overload, oi.cpp:138). Returns (gridpoints/s, kind, threads, seconds, n)."""
    from oracle import bindings as B
    kind = "reference" if B.available("ref") else "port"
    if kind == "port" and not B.available("oracle"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    lib = B.load("ref" if kind == "reference" else "oracle")
    threads = threads or os.cpu_count() or 1
    lib.set_omp_threads(threads)
    total = w["y"].size
    stride = max(1, total // n_sample)
    sel = np.arange(0, total, stride)[:n_sample]
    timing = []
    lib.optimal_interpolation((w["y"].ravel()[sel], w["x"].ravel()[sel], None, None), w["background"].ravel()[sel],
                              (w["py"], w["px"], None, None), w["pobs"], w["pratios"], w["pbackground"],
                              B.make_structure(B.BARNES, H_SCALE), MAX_POINTS, B.CARTESIAN, timing=timing)
    return sel.size / timing[0], kind, threads, timing[0], int(sel.size)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload()
    n = 200000
    gps, kind, threads, sec, n = cpu_reference_run(w, n)            # calibration / first warm-up
    full_pass = None
    if kind == "reference" and N_GRID * N_GRID / gps < 150.0 and not args.quick:
        # one pass over the WHOLE 4000 x 4000 grid through the Grid overload (oi.cpp:26-87: to_points() + two full-field
        # vec2 <-> vec copies inside the call), so that the subsampled steps below can be checked against it
        try:
            from oracle import bindings as B
            lib = B.load("ref")
            lib.set_omp_threads(threads)
            timing = []
            lib.optimal_interpolation_grid(w["y"], w["x"], w["background"], (w["py"], w["px"]), w["pobs"], w["pratios"], w["pbackground"],
                                           B.make_structure(B.BARNES, H_SCALE), MAX_POINTS, B.CARTESIAN, timing=timing)
            full_pass = {"seconds": timing[0], "gridpoints/s": N_GRID * N_GRID / timing[0], "overload": "Grid (oi.cpp:26-87)", "threads": threads}
        except Exception as e:
            full_pass = {"error": repr(e)}
    n = int(min(4_000_000, max(50_000, gps * 8.0)))                  # ~8 s per step
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_run(w, n)
    t = []
    for _ in range(args.steps):
        gps, kind, threads, sec, n = cpu_reference_run(w, n)
        t.append(sec)
    value = n * len(t) / sum(t)
    sample = "%d row-strided gridpoints of the 4000x4000 grid per step (Points overload), %d host threads" % (n, threads)
    line = {"impl": "reference", "metric": "OI analysis gridpoints/sec", "value": value, "unit": "gridpoints/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / len(t), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "background": BACKGROUND, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "gridpoints/s", "cores": threads, "kind": kind, "sample": sample,
                             "full_grid_pass": full_pass},
            "e2e": {"value": value, "unit": "gridpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def _cpu_lib():
    from oracle import bindings as B
    kind = "reference" if B.available("ref") else "port"
    if kind == "port" and not B.available("oracle"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    lib = B.load("ref" if kind == "reference" else "oracle")
    threads = os.cpu_count() or 1
    lib.set_omp_threads(threads)
    return B, lib, kind, threads


def secondary_metrics(gpp, gd, torch, hbm_peak, with_cpu=True):
    """The other configurations of BASELINE.json on this GPU, device-resident, CUDA-event timed:
    config 2 (neighbourhood mean/min/max, 4000 x 4000, halfwidth 7) and the single-GPU form of config 4
    (neighbourhood_quantile_fast, halfwidth 15, 20 thresholds, 4000 x 4000 and 8000 x 8000) as GB/s of the 8 B/pixel
    algorithmic traffic and as a fraction of the measured HBM peak. Inputs rotate over buffers totalling >= 256 MB
    (> L2) so that every launch reads from HBM. Each entry carries the reference's CPU time for the same call at full
    size, all host threads, and the GPU result is checked against it on the way."""
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, bufs, reps):
        for i in range(3):
            fn(bufs[i % len(bufs)])
        ev0.record()
        for i in range(reps):
            fn(bufs[i % len(bufs)])
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    def entry(ms, n):
        gbs = 8.0 * n * n / (ms * 1e-3) / 1e9
        return {"ms": ms, "GB/s": gbs, "frac_of_hbm_peak": gbs / hbm_peak}

    def cpu(fn, n, got, exact):
        """fn(timing) runs the reference on the host and returns its result; compared with the GPU result `got`."""
        if not with_cpu:
            return None
        try:
            B, lib, kind, threads = _cpu_lib()
            timing = []
            want = fn(B, lib, timing)
            g = got.cpu().numpy()
            if exact:
                ok = bool(np.array_equal(g, want, equal_nan=True))
            else:
                ok = bool(np.array_equal(np.isnan(g), np.isnan(want)) and
                          np.nanmax(np.abs(g - want) / np.maximum(np.abs(want), 10.0)) <= 1e-5)
            return {"ms": 1e3 * timing[0], "GB/s": 8.0 * n * n / timing[0] / 1e9, "cores": threads, "kind": kind,
                    "sample": "full size", "gpu_equals_cpu": ok}
        except Exception as e:
            return {"error": repr(e)}

    n = N_GRID
    bufs = [torch.rand((n, n), device="cuda") * 10 for _ in range(4)]
    res = torch.empty((n, n), device="cuda")
    out["copy_4000x4000 (torch copy_, the practical ceiling at this size)"] = entry(timeit(lambda b: res.copy_(b), bufs, 20), n)
    host0 = bufs[0].cpu().numpy()
    for name, st, exact in (("mean", gpp.Mean, False), ("min", gpp.Min, True), ("max", gpp.Max, True)):
        e = entry(timeit(lambda b: gd.neighbourhood(b, 7, st, out=res), bufs, 20), n)
        gd.neighbourhood(bufs[0], 7, st, out=res)
        e["cpu_baseline"] = cpu(lambda B, lib, t: lib.neighbourhood(host0, 7, st, timing=t), n, res, exact)
        out["neighbourhood_%s_hw7_4000x4000" % name] = e
    # a field with missing values: the column sums must carry validity counts (the slow path of the mean kernel)
    nan_bufs = []
    for b in bufs:
        c = b.clone()
        c[torch.rand((n, n), device="cuda") < 0.01] = float("nan")
        nan_bufs.append(c)
    out["neighbourhood_mean_hw7_4000x4000_1pct_missing"] = entry(timeit(lambda b: gd.neighbourhood(b, 7, gpp.Mean, out=res), nan_bufs, 20), n)
    del nan_bufs
    thr = np.linspace(0, 10, 20).astype(np.float32)
    out["quantile_fast_hw15_T20_4000x4000"] = entry(timeit(lambda b: gd.neighbourhood_quantile_fast(b, 0.5, 15, thr, out=res), bufs, 6), n)
    del bufs, res
    n = 8000
    bufs = [torch.rand((n, n), device="cuda") * 10 for _ in range(2)]
    res = torch.empty((n, n), device="cuda")
    for name, st in (("mean", gpp.Mean), ("max", gpp.Max)):
        out["neighbourhood_%s_hw7_8000x8000" % name] = entry(timeit(lambda b: gd.neighbourhood(b, 7, st, out=res), bufs, 10), n)
    e = entry(timeit(lambda b: gd.neighbourhood_quantile_fast(b, 0.5, 15, thr, out=res), bufs, 4), n)
    gd.neighbourhood_quantile_fast(bufs[0], 0.5, 15, thr, out=res)
    host0 = bufs[0].cpu().numpy()
    e["cpu_baseline"] = cpu(lambda B, lib, t: lib.neighbourhood_quantile_fast(host0, 0.5, 15, thr, timing=t), n, res, True)
    out["quantile_fast_hw15_T20_8000x8000 (config 4 on one GPU)"] = e
    return out


ENSI_N, ENSI_DX, ENSI_E, ENSI_S, ENSI_MP = 2500, 200.0, 20, 5000, 50
ENSI_WORKLOAD = "C5: optimal_interpolation_ensi 2500x2500 grid (dx 200 m), 20 members, 5000 obs, BarnesStructure(10000), max_points 50"


def ensi_inputs(row0, row1):
    """Config 5 inputs for rows [row0, row1) of the grid; the ensemble is generated row by row from per-row seeds so that
    any rank can produce any rows (and the whole field is the same whoever generates it)."""
    n, dx, E, S = ENSI_N, ENSI_DX, ENSI_E, ENSI_S
    rng = np.random.default_rng(SEED)
    y, x = np.meshgrid(np.arange(row0, row1, dtype=np.float32) * dx, np.arange(n, dtype=np.float32) * dx, indexing="ij")
    py, px = (rng.random(S) * n * dx).astype(np.float32), (rng.random(S) * n * dx).astype(np.float32)
    pbg = rng.standard_normal((S, E)).astype(np.float32)
    obs = rng.standard_normal(S).astype(np.float32)
    sig = np.full(S, 0.5, np.float32)
    bg = np.empty((row1 - row0, n, E), np.float32)
    for r in range(row0, row1):
        rr = np.random.default_rng(SEED * 100003 + r)
        bg[r - row0] = rr.standard_normal((n, E), dtype=np.float32) + 2 * rr.standard_normal((n, 1), dtype=np.float32)
    return dict(y=y, x=x, py=py, px=px, pbg=pbg, obs=obs, sig=sig, bg=bg)


def ensi_flops_per_gridpoint(E=ENSI_E, k=ENSI_MP):
    """2 (2 E^2 k + E k) + F_eig(E) + 4 E^3 + 2 E^2 with F_eig = 36 E^3 (cyclic Jacobi, ~6 sweeps)."""
    return 2.0 * (2 * E * E * k + E * k) + 36.0 * E ** 3 + 4.0 * E ** 3 + 2.0 * E * E


def ensi_metric(gpp, gd, torch, fp64_peak=None, with_cpu=True):
    """Config 5 on one GPU: (1) the kernel alone, device-resident, CUDA-event timed (gpp_optimal_interpolation_ensi_device)
    with its fp64 roofline; (2) end to end through the host API (H2D of the 500 MB ensemble, kernel, D2H); (3) the
    reference's serial loop (oi_ensi.cpp:203-207) on a row-strided sample of the same points, checked against the GPU."""
    w = ensi_inputs(0, ENSI_N)
    n, E = ENSI_N, ENSI_E
    grid, points = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian), gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
    s = gpp.BarnesStructure(H_SCALE)
    out = {"workload": ENSI_WORKLOAD}
    d_bg = torch.from_numpy(w["bg"]).cuda()
    d_out = torch.empty_like(d_bg)
    state = gd.EnsembleObservationState(points, w["obs"], w["sig"], w["pbg"], s)
    gd.optimal_interpolation_ensi(grid, d_bg, state, ENSI_MP, out=d_out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(2):
        gd.optimal_interpolation_ensi(grid, d_bg, state, ENSI_MP, out=d_out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    kernel_ms = min(ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
    out["kernel_ms"] = kernel_ms
    out["kernel_gridpoints/s"] = n * n / (kernel_ms * 1e-3)
    f_gp = ensi_flops_per_gridpoint()
    tf = f_gp * n * n / (kernel_ms * 1e-3) / 1e12
    out["roofline_fp64"] = {"bound": "fp64_fma", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": tf / fp64_peak if fp64_peak else None, "flops_per_gridpoint": f_gp,
                            "hbm_bytes_per_gridpoint": 8 * E + 8, "kernel": "ensi_kernel"}
    dev = d_out.cpu().numpy()
    del d_bg, d_out, state
    t = []
    for _ in range(2):
        t0 = time.perf_counter()
        host = gpp.optimal_interpolation_ensi(grid, w["bg"], points, w["obs"], w["sig"], w["pbg"], s, ENSI_MP)
        t.append(time.perf_counter() - t0)
    out["seconds_end_to_end"] = t[-1]
    out["gridpoints/s"] = n * n / t[-1]
    out["h2d_bytes"] = int(w["bg"].nbytes)
    out["d2h_bytes"] = int(w["bg"].nbytes)
    out["device_and_host_entry_points_identical"] = bool(np.array_equal(dev, host))
    if with_cpu:
        try:
            B, lib, kind, threads = _cpu_lib()
            structure = B.make_structure(B.BARNES, H_SCALE)

            def run(m):
                pick = np.arange(0, n * n, max(1, n * n // m))[:m]
                timing = []
                want = lib.optimal_interpolation_ensi((w["y"].ravel()[pick], w["x"].ravel()[pick], None, None), w["bg"].reshape(-1, E)[pick],
                                                      (w["py"], w["px"], None, None), w["obs"], w["sig"], w["pbg"], structure, ENSI_MP,
                                                      B.CARTESIAN, timing=timing)
                return pick, want, timing[0]
            pick, want, sec = run(4000)
            m = int(min(100_000, max(4000, 4000 / sec * 20.0)))    # ~20 s of serial CPU work, at most SURVEY 8d's 100k points
            pick, want, sec = run(m)
            err = np.abs(host.reshape(-1, E)[pick] - want) / np.maximum(np.abs(want), 2.0)
            out["cpu_baseline"] = {"value": m / sec, "unit": "gridpoints/s", "cores": 1, "kind": kind, "seconds": sec,
                                   "sample": "%d row-strided gridpoints of the same workload (Points overload; the reference's loop is serial, "
                                             "oi_ensi.cpp:203-207)" % m,
                                   "gpu_vs_cpu_max_rel_err": float(err.max())}
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    return out


def ensi_multi_metric(gpp):
    """SURVEY 8(f)#1 on config 5's geometry: optimal_interpolation_ensi_multi_ebesc (member-by-member increments, static
    correlations, max_points 30) end to end through the host API, and the reference's serial loop (oi_ensi_multi.cpp:715) on a
    row-strided sample of the same points, checked against the GPU result."""
    w = ensi_inputs(0, ENSI_N)
    n, E, mp = ENSI_N, ENSI_E, 30
    grid, points = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian), gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
    s = gpp.BarnesStructure(H_SCALE)
    pobs = (w["obs"][:, None] + 0.1 * np.arange(E, dtype=np.float32)[None, :]).astype(np.float32)
    pratios = np.full(ENSI_S, 0.25, np.float32)
    bratios = np.ones((n, n), np.float32)
    t = []
    for _ in range(2):
        t0 = time.perf_counter()
        host = gpp.optimal_interpolation_ensi_multi_ebesc(grid, bratios, w["bg"], points, pobs, pratios, w["pbg"], s, mp, False)
        t.append(time.perf_counter() - t0)
    out = {"workload": "optimal_interpolation_ensi_multi_ebesc %dx%d grid, %d members, %d obs, BarnesStructure(%g), max_points %d, clamp on"
                       % (n, n, E, ENSI_S, H_SCALE, mp),
           "seconds_end_to_end": t[-1], "gridpoints/s": n * n / t[-1], "h2d_bytes": int(w["bg"].nbytes + bratios.nbytes), "d2h_bytes": int(w["bg"].nbytes)}
    try:
        B, lib, kind, threads = _cpu_lib()
        structure = B.make_structure(B.BARNES, H_SCALE)

        def run(m):
            pick = np.arange(0, n * n, max(1, n * n // m))[:m]
            t0 = time.perf_counter()
            want = lib.ensi_multi("ebesc", (w["y"].ravel()[pick], w["x"].ravel()[pick], None, None), bratios.ravel()[pick], w["bg"].reshape(-1, E)[pick],
                                  None, (w["py"], w["px"], None, None), pobs, pratios, w["pbg"], None, structure, mp, B.CARTESIAN, False)
            return pick, want, time.perf_counter() - t0
        pick, want, sec = run(4000)
        m = int(min(200_000, max(4000, 4000 / sec * 15.0)))
        pick, want, sec = run(m)
        err = np.abs(host.reshape(-1, E)[pick] - want) / np.maximum(np.abs(want), 2.0)
        out["cpu_baseline"] = {"value": m / sec, "unit": "gridpoints/s", "cores": 1, "kind": kind, "seconds": sec,
                               "sample": "%d row-strided gridpoints of the same workload (Points overload; the reference's loop is serial)" % m,
                               "gpu_vs_cpu_max_rel_err": float(err.max())}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def ensi_sharded_metric(gpp, torch, dist, rank, world, barrier):
    """Config 5 row-sharded: every rank analyses its block of rows against the full observation set through the host API
    (grid points are independent, oi_ensi.cpp:207-554; no member is invalid here, so the mask of :187-201 needs no
    exchange); seconds = max over ranks. Afterwards every rank analyses the WHOLE field on its own GPU and compares its
    rows (to 1e-6: the Jacobi warm-start chain of a point depends on where its block of 32 points starts)."""
    r0, r1 = ENSI_N * rank // world, ENSI_N * (rank + 1) // world
    w = ensi_inputs(r0, r1)
    grid, points = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian), gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
    s = gpp.BarnesStructure(H_SCALE)
    t = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        mine = gpp.optimal_interpolation_ensi(grid, w["bg"], points, w["obs"], w["sig"], w["pbg"], s, ENSI_MP)
        t.append(time.perf_counter() - t0)
    sec = torch.tensor([t[-1]], device="cuda", dtype=torch.float64)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    full = ensi_inputs(0, ENSI_N)
    whole = gpp.optimal_interpolation_ensi(gpp.Grid(full["y"], full["x"], type=gpp.Cartesian), full["bg"], points, w["obs"], w["sig"],
                                           w["pbg"], s, ENSI_MP)
    same = bool(np.allclose(mine, whole[r0:r1], rtol=1e-6, atol=1e-6))
    flag = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"seconds_end_to_end": float(sec.item()), "gridpoints/s": ENSI_N * ENSI_N / float(sec.item()),
            "sharded_equals_whole": bool(flag.item()), "tolerance": 1e-6}


def halo_neighbourhood_metric(gpp, gd, torch, dist, world, rank, hbm_peak):
    """Config 4 across the ranks: an 8000 x 8000 field row-tiled over the GPUs, `halfwidth` halo rows from the vertical
    neighbours, then the filter on tile + halo. Timed per step: halo exchange + kernel(s), CUDA events, max over ranks.
    Every rank generates the whole field from the same seed, so each sharded result is also compared (bit for bit) with
    the rank's rows of the single-GPU whole-field result: `sharded_equals_whole`."""
    from gridpp_b200 import distributed as gdist
    n, hw = 8000, 15
    gen = torch.Generator(device="cuda")
    gen.manual_seed(SEED)
    field = torch.rand((n, n), device="cuda", generator=gen)
    thr = np.linspace(0, 1, 20).astype(np.float32)
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    modes = [("nccl", {})]
    if getattr(gdist, "peer_halo_available", lambda: False)():
        modes.append(("peer", {"halo": "peer"}))
    for mode, kw in modes:
        tile = gdist.RowTile(n, n, hw, device="cuda", **kw)          # this rank's rows, stored with room for both halos
        tile.tile.copy_(field[tile.r0:tile.r1])
        res = torch.empty((tile.rows, n), device="cuda")
        for name, fn, whole_fn, reps in (
                ("neighbourhood_mean_hw15", lambda: gdist.neighbourhood(tile, hw, gpp.Mean, out=res), lambda: gd.neighbourhood(field, hw, gpp.Mean), 10),
                ("quantile_fast_hw15_T20", lambda: gdist.neighbourhood_quantile_fast(tile, 0.5, hw, thr, out=res),
                 lambda: gd.neighbourhood_quantile_fast(field, 0.5, hw, thr), 5)):
            for _ in range(2):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1) / reps], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            whole = whole_fn()
            same = torch.equal(torch.nan_to_num(res, nan=-7.0), torch.nan_to_num(whole[tile.r0:tile.r1], nan=-7.0))
            flag = torch.tensor([int(same)], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            del whole
            ms = float(t.item())
            gbs = 8.0 * n * n / (ms * 1e-3) / 1e9
            out["%s_8000x8000_rows_over_%d_gpus_%s_halo" % (name, world, mode)] = {
                "ms": ms, "GB/s": gbs, "frac_of_aggregate_hbm_peak": gbs / (hbm_peak * world),
                "halo_bytes_per_rank_per_step": int(2 * hw * n * 4), "sharded_equals_whole": bool(flag.item())}
        del tile, res
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gridpp_b200 as gpp
    from gridpp_b200 import device as gd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if gpp.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    gpp.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # rows of this rank (contiguous block)
    rows = [N_GRID * r // world for r in range(world + 1)]
    row0, row1 = rows[rank], rows[rank + 1]
    w = make_workload(row0, row1)
    n_local = (row1 - row0) * N_GRID
    n_total = N_GRID * N_GRID

    grid = gpp.Grid(w["y"], w["x"], type=gpp.Cartesian)
    points = gpp.Points(w["py"], w["px"], type=gpp.Cartesian)
    structure = gpp.BarnesStructure(H_SCALE)
    state = gd.ObservationState(points, w["pobs"], w["pratios"], w["pbackground"], structure)
    d_bg = torch.from_numpy(w["background"].ravel()).cuda()
    d_out = torch.empty_like(d_bg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        gd.optimal_interpolation(grid, d_bg, state, MAX_POINTS, out=d_out)

    # clocks / throttle reasons are sampled from before the warm-up steps (the same load) to the end of the timed steps:
    # the timed region alone (10 x 39 ms) is too short for more than a couple of nvidia-smi samples
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = gpp.kernel_launch_count()
    # Timing rule: inputs larger than L2, or L2 flushed between timed iterations. One rank's inputs + outputs are 16 B x its
    # grid points: 256 MB at N = 1 (larger than the 126 MB L2), 64 / 32 MB at N = 4 / 8 -- there a 256 MB buffer is rewritten
    # between the steps, outside the per-step event pairs that are summed.
    flush_l2 = 16.0 * n_local < 1.5 * 126e6
    flush_buf = torch.empty(64 << 20, dtype=torch.float32, device="cuda") if flush_l2 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        if flush_l2:
            flush_buf.fill_(float(i))
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    launches = gpp.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    per_launch_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(per_launch_ms)
    del flush_buf
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n_total * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public host API, pinned host buffers
    h_bg = torch.from_numpy(w["background"]).pin_memory()
    bg_np = h_bg.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    out_np = gpp.optimal_interpolation(grid, bg_np, points, w["pobs"], w["pratios"], w["pbackground"], structure, MAX_POINTS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_np = gpp.optimal_interpolation(grid, bg_np, points, w["pobs"], w["pratios"], w["pbackground"], structure, MAX_POINTS)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(t.item())
    # the two entry points must agree bit for bit
    same = bool(np.array_equal(out_np.ravel(), d_out.cpu().numpy(), equal_nan=True))
    # N > 1: this rank's rows must equal the same rows of the whole 4000 x 4000 analysis on one GPU
    sharded_same = None
    if world > 1:
        wf = make_workload()
        whole = gd.optimal_interpolation(gpp.Grid(wf["y"], wf["x"], type=gpp.Cartesian), torch.from_numpy(wf["background"].ravel()).cuda(),
                                         state, MAX_POINTS)
        mine = whole[row0 * N_GRID:row1 * N_GRID]
        flag = torch.tensor([int(torch.equal(torch.nan_to_num(mine, nan=-7.0), torch.nan_to_num(d_out, nan=-7.0)))], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_same = bool(flag.item())
        del whole, mine, wf

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # dram__bytes_read.sum + dram__bytes_write.sum of one full-grid launch of the current kernel, from the committed ncu
    # capture (profiles/r2_traffic.json names the summary file it was read from)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["oi_fast_kernel"]
        traffic = float(tr["dram_bytes_per_full_grid_launch"]) * n_local / n_total
    except (OSError, ValueError, KeyError):
        pass
    halo = None
    if world > 1 and not args.quick:
        # every rank takes part: row tiles + halo rows for the stencil filters (config 4)
        try:
            halo = halo_neighbourhood_metric(gpp, gd, torch, dist, world, rank, hbm_peak)
        except Exception as e:
            halo = {"error": repr(e)}
        try:   # config 5 sharded by rows
            halo["ensi_2500x2500x20_rows_over_%d_gpus" % world] = ensi_sharded_metric(gpp, torch, dist, rank, world, barrier)
        except Exception as e:
            halo["ensi_error"] = repr(e)
    if rank == 0:
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        kernel_ms = statistics.mean(per_launch_ms)          # one kernel launch per step on this rank
        # bg 4 + analysis 4 + lat/lon 8 per gridpoint; the kernel reads exactly these planes for a Cartesian grid
        # without elevations / land fractions (z, elev, laf are not read)
        bytes_per_launch = 16.0 * n_local
        achieved_gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        config = {"workload": WORKLOAD, "background": BACKGROUND, "grid_rows_per_gpu": row1 - row0,
                  "parallelism": "rows split over %d GPU(s), observations replicated, no collective" % world,
                  "l2": "per-step inputs+outputs are %.0f MB per GPU (background, analysis, 2 coordinate planes); %s"
                        % (16.0 * n_local / 1e6, "L2 flushed between the timed steps (a 256 MB buffer rewritten outside the event pairs)" if flush_l2
                           else "larger than the 126 MB L2, no flush"),
                  "device_and_host_entry_points_identical": same}
        if sharded_same is not None:
            config["sharded_equals_whole"] = sharded_same
        line = {
            "metric": "OI analysis gridpoints/sec", "value": value, "unit": "gridpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "gridpoints/s", "h2d_bytes_per_step": int(4 * n_total + 12 * N_OBS),
                    "d2h_bytes_per_step": int(4 * n_total), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": traffic, "kernel": "oi_fast_kernel", "peak_source": peak_src,
                         "note": "HBM view only for the contract; the kernel is fp64-CUDA-core / issue bound, see roofline_fp64"},
        }
        fp64_peak = None
        if not args.quick:
            try:
                f_gp, nc_mean, k_mean = oi_flops_per_gridpoint(w if world == 1 else make_workload(0, 1))
                fp64_peak = gpp.measure_fp64_fma_peak()
                achieved_tf = f_gp * n_local / (kernel_ms * 1e-3) / 1e12
                line["roofline_fp64"] = {"bound": "fp64_fma", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                                         "frac": achieved_tf / fp64_peak, "flops_per_gridpoint": f_gp, "mean_candidates": nc_mean,
                                         "mean_selected": k_mean, "peak_source": "measured in this run (gpp_measure_fp64_fma_peak)"}
            except Exception as e:   # never lose the headline over an auxiliary figure
                line["roofline_fp64"] = {"error": repr(e)}
            if world == 1:
                try:
                    gps, kind, threads, sec, n = cpu_reference_run(w, 100000)
                    n2 = int(min(3_000_000, max(100_000, gps * 12.0)))
                    gps, kind, threads, sec, n = cpu_reference_run(w, n2)
                    line["cpu_baseline"] = {"value": gps, "unit": "gridpoints/s", "cores": threads, "kind": kind, "seconds": sec,
                                            "sample": "%d row-strided gridpoints of the same 4000x4000 workload (Points overload)" % n}
                except Exception as e:
                    line["cpu_baseline"] = {"error": repr(e)}
                try:
                    line["secondary"] = full_variance_metric(gpp, gd, torch, grid, state, d_bg)
                except Exception as e:
                    line["secondary"] = {"optimal_interpolation_full": {"error": repr(e)}}
                del d_bg, d_out
                try:
                    line["secondary"].update(secondary_metrics(gpp, gd, torch, hbm_peak))
                except Exception as e:
                    line["secondary"]["error"] = repr(e)
                try:
                    line["secondary"]["ensi"] = ensi_metric(gpp, gd, torch, fp64_peak)
                except Exception as e:
                    line["secondary"]["ensi"] = {"error": repr(e)}
                try:
                    line["secondary"]["ensi_multi_ebesc"] = ensi_multi_metric(gpp)
                except Exception as e:
                    line["secondary"]["ensi_multi_ebesc"] = {"error": repr(e)}
        if halo is not None:
            line["secondary_multi_gpu"] = halo
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def full_variance_metric(gpp, gd, torch, grid, state, d_bg):
    """optimal_interpolation_full on config 3: the analysis-variance output needs rho' (P+R)^-1 rho per grid point, which the
    plain analysis does not; device-resident, CUDA-event timed."""
    d_out, d_var = torch.empty_like(d_bg), torch.empty_like(d_bg)
    gd.optimal_interpolation(grid, d_bg, state, MAX_POINTS, out=d_out, out_variance=d_var)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(2):
        gd.optimal_interpolation(grid, d_bg, state, MAX_POINTS, out=d_out, out_variance=d_var)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 2
    ok = bool(torch.isfinite(d_var).all().item() and float(d_var.max()) <= 1.0 + 1e-5 and float(d_var.min()) >= 0.0)
    return {"optimal_interpolation_full_C3 (analysis + analysis variance)": {"ms": ms, "gridpoints/s": N_GRID * N_GRID / (ms * 1e-3),
                                                                             "variance_in_[0,1]": ok}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="skip cpu_baseline, fp64 roofline and secondary metrics")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
