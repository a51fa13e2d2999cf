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
