"""World-size-2 `gloo` tests (CPU) of the multi-GPU plumbing: row blocks, halo exchange, tile-with-halo evaluation.
The oracle stands in for the CUDA kernels (it is injected as the `compute` callable); the N > 1 device path itself is
exercised by bench.py --gpus N on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bindings as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_tile(field, hw, st, row0, n_rows_out):
    """What gpp_neighbourhood_device does with (n_rows_in, row0, n_rows_out): the rows given ARE the domain."""
    orc = B.load("oracle")
    full = orc.neighbourhood(field.numpy(), hw, st)
    return torch.from_numpy(full[row0:row0 + n_rows_out].copy())


def _oracle_tile_q(field, q, hw, thr, row0, n_rows_out):
    orc = B.load("oracle")
    full = orc.neighbourhood_quantile_fast(field.numpy(), q, hw, thr)
    return torch.from_numpy(full[row0:row0 + n_rows_out].copy())


def _worker(rank, world, port, ny, nx, hw, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    from gridpp_b200 import distributed as gdist
    rng = np.random.default_rng(1000)
    field = (rng.uniform(size=(ny, nx)) * 10).astype(np.float32)
    field[rng.uniform(size=field.shape) < 0.02] = np.nan
    r0, r1 = gdist.row_block(ny, world, rank)
    tile = torch.from_numpy(field[r0:r1].copy())
    ext, above = gdist.exchange_halo(tile, hw)
    lo, hi = max(0, r0 - hw), min(ny, r1 + hw)
    ok = np.array_equal(ext.numpy(), field[lo:hi], equal_nan=True) and above == r0 - lo
    outs = {}
    for name, st in (("mean", B.MEAN), ("min", B.MIN), ("count", B.COUNT)):
        mine = gdist.neighbourhood(tile, hw, st, compute=_oracle_tile)
        outs[name] = gdist.gather_rows(mine).numpy()
    thr = np.linspace(0, 10, 9).astype(np.float32)
    mine = gdist.neighbourhood_quantile_fast(tile, 0.5, hw, thr, compute=_oracle_tile_q)
    outs["qfast"] = gdist.gather_rows(mine).numpy()
    np.savez(os.path.join(result_dir, "rank%d.npz" % rank), ok=ok, **outs)
    dist.destroy_process_group()


@pytest.mark.parametrize("ny,nx,hw", [(37, 23, 3), (64, 40, 7), (20, 16, 7)])
def test_row_sharded_filters_equal_whole(tmp_path, orc, ny, nx, hw):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ny, nx, hw, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(1000)
    field = (rng.uniform(size=(ny, nx)) * 10).astype(np.float32)
    field[rng.uniform(size=field.shape) < 0.02] = np.nan
    want = {"mean": orc.neighbourhood(field, hw, B.MEAN), "min": orc.neighbourhood(field, hw, B.MIN),
            "count": orc.neighbourhood(field, hw, B.COUNT),
            "qfast": orc.neighbourhood_quantile_fast(field, 0.5, hw, np.linspace(0, 10, 9).astype(np.float32))}
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert bool(got["ok"]), "halo exchange delivered the wrong rows on rank %d" % rank
        for name, w in want.items():
            assert np.array_equal(got[name], w, equal_nan=True), (rank, name)


def test_row_block_partition():
    import sys
    sys.path.insert(0, ROOT)
    from gridpp_b200 import distributed as gdist
    for n, world in ((4000, 8), (2500, 8), (7, 3), (5, 8)):
        blocks = [gdist.row_block(n, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _oi_inputs():
    rng = np.random.default_rng(5)
    ny, nx, S, E = 22, 17, 60, 5
    y, x = np.meshgrid(np.arange(ny) * 1000.0, np.arange(nx) * 1000.0, indexing="ij")
    y, x = y.astype(np.float32), x.astype(np.float32)
    py, px = rng.uniform(0, ny * 1000, S).astype(np.float32), rng.uniform(0, nx * 1000, S).astype(np.float32)
    bg = rng.normal(size=(ny, nx)).astype(np.float32)
    ens = (bg[:, :, None] + rng.normal(size=(ny, nx, E))).astype(np.float32)
    ens[3, 4, 2] = np.nan                      # lives in rank 0's rows only: rank 1 must learn about it
    pbg, pens = rng.normal(size=S).astype(np.float32), rng.normal(size=(S, E)).astype(np.float32)
    obs = (pbg + rng.normal(size=S)).astype(np.float32)
    return dict(y=y, x=x, py=py, px=px, bg=bg, ens=ens, pbg=pbg, pens=pens, obs=obs, ratios=np.full(S, 0.5, np.float32),
                sig=np.full(S, 0.7, np.float32))


def _oi_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    from gridpp_b200 import distributed as gdist
    d = _oi_inputs()
    orc = B.load("oracle")
    s = B.make_structure(B.BARNES, 5000.0)

    def oi_rows(lats, lons, elevs, lafs, background):
        out = orc.optimal_interpolation((lats, lons, None, None), background, (d["py"], d["px"], None, None), d["obs"], d["ratios"],
                                        d["pbg"], s, 10, B.CARTESIAN)
        return out.reshape(background.shape)

    def ensi_rows(lats, lons, elevs, lafs, background):
        out = orc.optimal_interpolation_ensi((lats, lons, None, None), background.reshape(-1, background.shape[-1]),
                                             (d["py"], d["px"], None, None), d["obs"], d["sig"], d["pens"], s, 10, B.CARTESIAN)
        return out.reshape(background.shape)

    oi = gdist.optimal_interpolation(d["y"], d["x"], d["bg"], None, d["obs"], d["ratios"], d["pbg"], None, 10, type=1, gather=True, compute=oi_rows)
    ensi = gdist.optimal_interpolation_ensi(d["y"], d["x"], d["ens"], None, d["obs"], d["sig"], d["pens"], None, 10, type=1, gather=True,
                                            compute=ensi_rows)
    np.savez(os.path.join(result_dir, "oi_rank%d.npz" % rank), oi=oi, ensi=ensi)
    dist.destroy_process_group()


def test_row_sharded_oi_and_ensi_equal_whole(tmp_path, orc):
    """OI and EnSI sharded by rows over two ranks (the oracle standing in for the kernels) equal the whole-field result,
    including the EnSI rule that a member with an invalid value ANYWHERE is left untouched on every rank."""
    world = 2
    mp.spawn(_oi_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    d = _oi_inputs()
    s = B.make_structure(B.BARNES, 5000.0)
    want_oi = orc.optimal_interpolation((d["y"], d["x"], None, None), d["bg"], (d["py"], d["px"], None, None), d["obs"], d["ratios"], d["pbg"],
                                        s, 10, B.CARTESIAN).reshape(d["bg"].shape)
    E = d["ens"].shape[-1]
    want_ensi = orc.optimal_interpolation_ensi((d["y"], d["x"], None, None), d["ens"].reshape(-1, E), (d["py"], d["px"], None, None), d["obs"],
                                               d["sig"], d["pens"], s, 10, B.CARTESIAN).reshape(d["ens"].shape)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "oi_rank%d.npz" % rank))
        assert np.array_equal(got["oi"], want_oi, equal_nan=True), rank
        assert np.array_equal(got["ensi"], want_ensi, equal_nan=True), rank
        assert np.array_equal(got["ensi"][:, :, 2], d["ens"][:, :, 2], equal_nan=True)   # the member with the missing value
