"""Multi-GPU parity worker: run under torchrun (one rank per GPU, NCCL), by tests/test_gpu_multi.py or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py

Every row-sharded driver of gridpp_b200.distributed is compared with the single-GPU whole-field result of the same
inputs (every rank builds the same field from the seed and evaluates the whole field on its own GPU):
  * neighbourhood Mean / Min / Max / Count and neighbourhood_quantile_fast over row tiles with a halo exchange
    (NCCL send/recv, and the peer-memory path when it is available) -- bit-equal;
  * optimal_interpolation over row blocks -- bit-equal; optimal_interpolation_ensi -- to 1e-6 (see below).
Rank 0 prints one JSON line {"n_gpus": N, "checks": {name: bool}}; the exit code is 1 when any check failed.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import device as gd, distributed as gdist


def same(a, b):
    return bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    gpp.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    checks = {}

    # ---- stencil filters: row tiles + halo rows
    ny, nx = 1800 + 7 * world, 2048
    g = torch.Generator(device="cuda").manual_seed(1000)
    field = torch.rand((ny, nx), device="cuda", generator=g) * 10
    field[torch.rand((ny, nx), device="cuda", generator=g) < 0.01] = float("nan")
    r0, r1 = gdist.row_block(ny, world, rank)
    tile = field[r0:r1].contiguous()
    thr = np.linspace(0, 10, 20).astype(np.float32)
    modes = ["nccl"] + (["peer"] if getattr(gdist, "peer_halo_available", lambda: False)() else [])
    for mode in modes:
        kw = {} if mode == "nccl" else {"halo": "peer"}
        for hw in (7, 15):
            for name, st in (("mean", gpp.Mean), ("min", gpp.Min), ("max", gpp.Max), ("count", gpp.Count)):
                mine = gdist.neighbourhood(tile, hw, st, **kw)
                whole = gd.neighbourhood(field, hw, st)
                checks["%s neighbourhood %s hw%d" % (mode, name, hw)] = same(mine, whole[r0:r1])
            mine = gdist.neighbourhood_quantile_fast(tile, 0.5, hw, thr, **kw)
            whole = gd.neighbourhood_quantile_fast(field, 0.5, hw, thr)
            checks["%s quantile_fast hw%d" % (mode, hw)] = same(mine, whole[r0:r1])

    # ---- OI / EnSI: row blocks, observations replicated
    rng = np.random.default_rng(1000)
    gy, gx, dx, S, E = 96 * world + 10, 400, 250.0, 900, 20
    y, x = np.meshgrid(np.arange(gy, dtype=np.float32) * dx, np.arange(gx, dtype=np.float32) * dx, indexing="ij")
    py, px = (rng.random(S) * gy * dx).astype(np.float32), (rng.random(S) * gx * dx).astype(np.float32)
    bg = rng.standard_normal((gy, gx)).astype(np.float32) * 3
    pbg = rng.standard_normal(S).astype(np.float32)
    obs = (pbg + rng.standard_normal(S) * 0.5).astype(np.float32)
    ratios = np.full(S, 0.5, np.float32)
    points, s = gpp.Points(py, px, type=gpp.Cartesian), gpp.BarnesStructure(10000)
    whole = gpp.optimal_interpolation(gpp.Grid(y, x, type=gpp.Cartesian), bg, points, obs, ratios, pbg, s, 30)
    mine = gdist.optimal_interpolation(y, x, bg, points, obs, ratios, pbg, s, 30, type=gpp.Cartesian)
    b0, b1 = gdist.row_block(gy, world, rank)
    checks["optimal_interpolation rows"] = bool(np.array_equal(mine, whole[b0:b1], equal_nan=True))
    gathered = gdist.optimal_interpolation(y, x, bg, points, obs, ratios, pbg, s, 30, type=gpp.Cartesian, gather=True)
    checks["optimal_interpolation gather"] = bool(np.array_equal(gathered, whole, equal_nan=True))
    ebg = (rng.standard_normal((gy, gx, 1)) * 2 + rng.standard_normal((gy, gx, E))).astype(np.float32)
    ebg[gy - 3, 5, 4] = np.nan     # one member is invalid on the LAST rank only: the flags must be all-reduced
    epbg = rng.standard_normal((S, E)).astype(np.float32)
    sig = np.full(S, 0.5, np.float32)
    ewhole = gpp.optimal_interpolation_ensi(gpp.Grid(y, x, type=gpp.Cartesian), ebg, points, obs, sig, epbg, s, 50)
    emine = gdist.optimal_interpolation_ensi(y, x, ebg, points, obs, sig, epbg, s, 50, type=gpp.Cartesian)
    # (the Jacobi warm-start chain runs inside blocks of 32 consecutive points of the rank's own grid, so a sharded point may
    # start its iteration elsewhere than the whole-field one: equal to the 1e-13 convergence level, not bit for bit)
    keep = [e for e in range(E) if e != 4]
    checks["optimal_interpolation_ensi rows"] = bool(np.allclose(emine[..., keep], ewhole[b0:b1][..., keep], rtol=1e-6, atol=1e-6))
    checks["optimal_interpolation_ensi invalid member untouched"] = bool(np.array_equal(emine[..., 4], ebg[b0:b1][..., 4], equal_nan=True))

    flags = torch.tensor([int(v) for v in checks.values()], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = {k: bool(f) for k, f in zip(checks.keys(), flags.tolist())}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "checks": ok}))
        sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(ok.values()) else 1)


if __name__ == "__main__":
    main()
