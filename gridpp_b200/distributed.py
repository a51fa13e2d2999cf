"""Row-sharded multi-GPU drivers: one process per GPU, ``torch.distributed`` for the plumbing.

* OI / EnSI: every output point is independent (oi.cpp:221-338), so each rank analyses a contiguous block of rows
  against the full (replicated, ~0.4 MB) observation table. No data-path collective.
* Neighbourhood filters: a stencil of radius ``halfwidth``; each rank needs ``halfwidth`` rows from its upper and
  lower neighbour (true domain edges are clipped, not padded: neighbourhood.cpp:104-107). ``exchange_halo`` moves
  those rows with point-to-point sends (NCCL over NVLink for CUDA tensors, gloo for CPU tensors in the tests);
  the filter then runs on the tile-with-halo through the ``*_device`` entry points with ``row0`` / ``n_rows_out``.

The compute callable defaults to the CUDA device entry point; tests/test_distributed_cpu.py injects a stand-in so
that the exchange plumbing itself can be exercised with gloo on a box without a GPU.
"""
import torch
import torch.distributed as dist


def row_block(n_rows, world_size, rank):
    """Contiguous, balanced block of rows [begin, end) owned by `rank`."""
    return n_rows * rank // world_size, n_rows * (rank + 1) // world_size


def exchange_halo(tile, halfwidth, group=None):
    """tile: (rows_local, nx) tensor holding this rank's rows. Returns (tile_with_halo, halo_above) where up to
    `halfwidth` rows of the previous / next rank have been attached above / below. Ranks own consecutive row blocks
    in rank order; a rank with fewer than `halfwidth` rows forwards what it has (the halo is then shorter, which is
    only correct when every block has at least `halfwidth` rows -- checked)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1 or halfwidth == 0:
        return tile, 0
    rows, nx = tile.shape
    n = torch.tensor([rows], dtype=torch.int64, device=tile.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    if min(sizes) < halfwidth:
        raise ValueError("every rank needs at least halfwidth=%d rows for a single-hop halo exchange (got %s)" % (halfwidth, sizes))
    up, down = rank - 1, rank + 1
    ops, recv_up, recv_down = [], None, None
    if up >= 0:
        recv_up = torch.empty((halfwidth, nx), dtype=tile.dtype, device=tile.device)
        ops.append(dist.P2POp(dist.isend, tile[:halfwidth].contiguous(), up, group))
        ops.append(dist.P2POp(dist.irecv, recv_up, up, group))
    if down < world:
        recv_down = torch.empty((halfwidth, nx), dtype=tile.dtype, device=tile.device)
        ops.append(dist.P2POp(dist.isend, tile[rows - halfwidth:].contiguous(), down, group))
        ops.append(dist.P2POp(dist.irecv, recv_down, down, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [p for p in (recv_up, tile, recv_down) if p is not None]
    return torch.cat(parts, dim=0), (halfwidth if recv_up is not None else 0)


def neighbourhood(tile, halfwidth, statistic, compute=None, group=None):
    """Row-sharded gridpp.neighbourhood: `tile` holds this rank's rows; returns this rank's rows of the result.
    `compute(field_with_halo, halfwidth, statistic, row0, n_rows_out)` defaults to the CUDA device entry point."""
    if compute is None:
        from . import device as gd

        def compute(field, hw, st, row0, n_rows_out):
            return gd.neighbourhood(field, hw, st, row0=row0, n_rows_out=n_rows_out)
    ext, above = exchange_halo(tile, halfwidth, group)
    return compute(ext.contiguous(), halfwidth, statistic, above, tile.shape[0])


def neighbourhood_quantile_fast(tile, quantile, halfwidth, thresholds, compute=None, group=None):
    """Row-sharded gridpp.neighbourhood_quantile_fast (scalar quantile)."""
    if compute is None:
        from . import device as gd

        def compute(field, q, hw, thr, row0, n_rows_out):
            return gd.neighbourhood_quantile_fast(field, q, hw, thr, row0=row0, n_rows_out=n_rows_out)
    ext, above = exchange_halo(tile, halfwidth, group)
    return compute(ext.contiguous(), quantile, halfwidth, thresholds, above, tile.shape[0])


def gather_rows(tile, group=None):
    """All-gather of the row blocks into the full field on every rank (optional; OI and the filters never need it)."""
    world = dist.get_world_size(group)
    if world == 1:
        return tile
    n = torch.tensor([tile.shape[0]], dtype=torch.int64, device=tile.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    padded = torch.zeros((m, tile.shape[1]), dtype=tile.dtype, device=tile.device)
    padded[:tile.shape[0]] = tile
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
