"""Row-sharded multi-GPU drivers: one process per GPU, ``torch.distributed`` for the plumbing.

* OI / EnSI: every output point is independent (oi.cpp:221-338), so each rank analyses a contiguous block of rows
  against the full (replicated, ~0.4 MB) observation table. No data-path collective: `optimal_interpolation` /
  `optimal_interpolation_ensi` below build the rank's Grid from its rows and call the ordinary entry points; only
  `gather=True` (every rank wants the whole field) adds an all-gather of the results.
* Neighbourhood filters: a stencil of radius ``halfwidth``; each rank needs ``halfwidth`` rows from its upper and
  lower neighbour (true domain edges are clipped, not padded: neighbourhood.cpp:104-107). A ``RowTile`` keeps the
  rank's rows inside a buffer with room for both halos, so the exchange (point-to-point, NCCL over NVLink for CUDA
  tensors, gloo for CPU tensors in the tests) lands in place and nothing is concatenated or re-allocated per step.
  The filter on the rows that need no halo is launched first and overlaps the exchange; the ``halfwidth`` rows at
  each end follow once the halo is there. All three launches go through the ``*_device`` entry points with
  ``row0`` / ``n_rows_out``, which evaluate a tile-with-halo exactly like the whole field.

The compute callable defaults to the CUDA device entry point; tests/test_distributed_cpu.py injects a stand-in so
that the exchange plumbing itself can be exercised with gloo on a box without a GPU.
"""
import torch
import torch.distributed as dist


def row_block(n_rows, world_size, rank):
    """Contiguous, balanced block of rows [begin, end) owned by `rank`."""
    return n_rows * rank // world_size, n_rows * (rank + 1) // world_size


def peer_halo_available():
    """True when the halo rows can be read straight out of the neighbours' memory: NCCL process group on CUDA devices and
    torch's symmetric memory (the allocation + handle exchange; the copy itself is gpp_halo_pull_device)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl" and torch.cuda.is_available()):
        return False
    try:
        import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
    except Exception:
        return False
    return True


class RowTile:
    """This rank's rows of a row-sharded (n_rows_global, nx) field, stored with space for `halfwidth` halo rows on
    each side. `tile` is the view to fill with the rank's own rows (row_block(n_rows_global, world, rank)).

    halo="nccl": the halo rows travel as a batched isend / irecv with the two vertical neighbours (gloo on CPU tensors in
    the tests). halo="peer": the buffer lives in symmetric memory, so every rank has its neighbours' buffers mapped;
    `exchange()` is then a barrier on the signal pad followed by ONE kernel (gpp_halo_pull_device) that reads the 2 x
    halfwidth boundary rows over NVLink -- no NCCL launch on the data path."""

    def __init__(self, n_rows_global, nx, halfwidth, device=None, dtype=torch.float32, group=None, halo="nccl"):
        self.group = group
        self.halo = halo
        self._symm = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.hw = int(halfwidth)
        self.n_rows_global, self.nx = int(n_rows_global), int(nx)
        self.r0, self.r1 = row_block(self.n_rows_global, self.world, self.rank)
        self.rows = self.r1 - self.r0
        if self.world > 1 and self.hw > 0:
            smallest = min(row_block(self.n_rows_global, self.world, r)[1] - row_block(self.n_rows_global, self.world, r)[0]
                           for r in range(self.world))
            if smallest < self.hw:
                raise ValueError("every rank needs at least halfwidth=%d rows for a single-hop halo exchange (smallest block: %d)"
                                 % (self.hw, smallest))
        self.above = self.hw if self.rank > 0 else 0                      # halo rows present above / below the tile
        self.below = self.hw if self.rank < self.world - 1 else 0
        if halo == "peer":
            import torch.distributed._symmetric_memory as symm_mem
            # symmetric allocations have the same size on every rank: room for the largest block
            self._rows_of = [row_block(self.n_rows_global, self.world, r)[1] - row_block(self.n_rows_global, self.world, r)[0]
                             for r in range(self.world)]
            whole = symm_mem.empty((max(self._rows_of) + 2 * self.hw, self.nx), dtype=dtype, device=device)
            self._symm = symm_mem.rendezvous(whole, group if group is not None else dist.group.WORLD)
            self._whole = whole
            self.buf = whole[:self.rows + 2 * self.hw]
        else:
            self.buf = torch.empty((self.rows + 2 * self.hw, self.nx), dtype=dtype, device=device)
        self.tile = self.buf[self.hw:self.hw + self.rows]

    @property
    def with_halo(self):
        """The tile with the halo rows that exist (domain edges have none)."""
        return self.buf[self.hw - self.above:self.hw + self.rows + self.below]

    def exchange_async(self):
        """Starts the halo exchange with the vertical neighbours; returns the requests to wait on."""
        if self.world == 1 or self.hw == 0:
            return []
        hw, ops = self.hw, []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self.tile[:hw], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[:hw], self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, self.tile[self.rows - hw:], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[hw + self.rows:], self.rank + 1, self.group))
        return dist.batch_isend_irecv(ops)

    def exchange(self):
        if self.halo == "peer":
            return self.exchange_peer()
        for req in self.exchange_async():
            req.wait()

    def exchange_peer(self):
        """Barrier (every rank's tile is complete, every earlier pull is done), then one kernel that copies the boundary rows
        out of the neighbours' buffers. Asynchronous on the current stream."""
        import ctypes as C
        from ._lib import check, lib
        if self.world == 1 or self.hw == 0:
            return
        esize = self.buf.element_size()
        row_bytes = self.nx * esize
        above = below = None
        if self.rank > 0:      # the last hw tile rows of rank - 1: its buffer rows [rows_up, rows_up + hw)
            above = C.c_void_p(int(self._symm.buffer_ptrs[self.rank - 1]) + self._rows_of[self.rank - 1] * row_bytes)
        if self.rank < self.world - 1:   # the first hw tile rows of rank + 1: its buffer rows [hw, 2 hw)
            below = C.c_void_p(int(self._symm.buffer_ptrs[self.rank + 1]) + self.hw * row_bytes)
        self._symm.barrier(channel=0)
        check(lib.gpp_halo_pull_device(C.c_void_p(self.buf.data_ptr()), self.rows, self.nx, self.hw, above, below,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))


def _run_tiled(rt, out, launch):
    """launch(field, row0, n_rows_out, out_rows): evaluates output rows [row0, row0 + n_rows_out) of `field` (whose
    first and last rows are treated as domain edges) into `out_rows`. The rows whose windows stay inside the tile go
    first and overlap the exchange; the (at most `halfwidth`) rows at each end follow once the halo is there."""
    hw, rows = rt.hw, rt.rows
    if rt.halo == "peer":
        # the pull is one short kernel on the same stream: one launch over the tile with its halo follows it
        rt.exchange_peer()
        launch(rt.with_halo, rt.above, rows, out)
        return out
    reqs = rt.exchange_async()
    if not reqs:
        launch(rt.with_halo, rt.above, rows, out)
        return out
    top = min(rows, hw if rt.above else 0)             # tile rows [0, top) need the upper halo
    bot_first = max(top, rows - (hw if rt.below else 0))   # tile rows [bot_first, rows) need the lower halo
    if bot_first > top:
        launch(rt.tile, top, bot_first - top, out[top:bot_first])
    for req in reqs:
        req.wait()
    full = rt.with_halo                                # tile row t sits at full[rt.above + t]
    if top > 0:
        end = min(rt.above + top + hw, full.shape[0])  # windows reach at most hw rows below the last of these rows
        launch(full[:end], rt.above, top, out[:top])
    if rows > bot_first:
        s0 = max(0, rt.above + bot_first - hw)         # ... and at most hw rows above the first of these
        launch(full[s0:], rt.above + bot_first - s0, rows - bot_first, out[bot_first:])
    return out


def neighbourhood(tile, halfwidth, statistic, compute=None, group=None, n_rows_global=None, out=None, halo="nccl"):
    """Row-sharded gridpp.neighbourhood. `tile` is either a RowTile (no copies) or a (rows_local, nx) tensor holding
    this rank's rows (copied into a RowTile; ranks own row_block() blocks). Returns this rank's rows of the result.
    `compute(field_with_halo, halfwidth, statistic, row0, n_rows_out)` defaults to the CUDA device entry point."""
    rt = _as_row_tile(tile, halfwidth, group, n_rows_global, halo)
    if out is None:
        out = torch.empty((rt.rows, rt.nx), dtype=rt.buf.dtype, device=rt.buf.device)
    if compute is None:
        from . import device as gd

        def launch(field, row0, n_rows_out, out_rows):
            gd.neighbourhood(field, halfwidth, statistic, out=out_rows, row0=row0, n_rows_out=n_rows_out)
    else:
        def launch(field, row0, n_rows_out, out_rows):
            out_rows.copy_(compute(field.contiguous(), halfwidth, statistic, row0, n_rows_out))
    return _run_tiled(rt, out, launch)


def neighbourhood_quantile_fast(tile, quantile, halfwidth, thresholds, compute=None, group=None, n_rows_global=None, out=None, halo="nccl"):
    """Row-sharded gridpp.neighbourhood_quantile_fast (scalar quantile)."""
    rt = _as_row_tile(tile, halfwidth, group, n_rows_global, halo)
    if out is None:
        out = torch.empty((rt.rows, rt.nx), dtype=rt.buf.dtype, device=rt.buf.device)
    if compute is None:
        from . import device as gd

        def launch(field, row0, n_rows_out, out_rows):
            gd.neighbourhood_quantile_fast(field, quantile, halfwidth, thresholds, out=out_rows, row0=row0, n_rows_out=n_rows_out)
    else:
        def launch(field, row0, n_rows_out, out_rows):
            out_rows.copy_(compute(field.contiguous(), quantile, halfwidth, thresholds, row0, n_rows_out))
    return _run_tiled(rt, out, launch)


def _as_row_tile(tile, halfwidth, group, n_rows_global, halo="nccl"):
    if isinstance(tile, RowTile):
        if tile.hw != int(halfwidth):
            raise ValueError("the RowTile was built for halfwidth %d" % tile.hw)
        return tile
    rows, nx = tile.shape
    if n_rows_global is None:
        n = torch.tensor([rows], dtype=torch.int64, device=tile.device)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(n, group=group)
        n_rows_global = int(n.item())
    rt = RowTile(n_rows_global, nx, halfwidth, device=tile.device, dtype=tile.dtype, group=group, halo=halo)
    if rt.rows != rows:
        raise ValueError("this rank holds %d rows, row_block() assigns it %d" % (rows, rt.rows))
    rt.tile.copy_(tile)
    return rt


def exchange_halo(tile, halfwidth, group=None):
    """Functional form: returns (tile_with_halo, halo_rows_above) for a (rows_local, nx) tensor."""
    rt = _as_row_tile(tile, halfwidth, group, None, "nccl")
    rt.exchange()
    return rt.with_halo, rt.above


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


def _rows_of(a, r0, r1):
    return None if a is None else a[r0:r1]


def _gather_result(rows, group):
    """All-gather of this rank's result rows (a numpy array) into the whole field, on the backend's device."""
    import numpy as np
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rows
    t = torch.from_numpy(np.ascontiguousarray(rows))
    flat = t.reshape(t.shape[0], -1)
    if dist.get_backend(group) == "nccl":
        flat = flat.cuda()
    return gather_rows(flat, group).cpu().numpy().reshape((-1,) + tuple(rows.shape[1:]))


def optimal_interpolation(lats, lons, background, points, pobs, pratios, pbackground, structure, max_points,
                          allow_extrapolation=True, elevs=None, lafs=None, type=0, group=None, gather=False, compute=None):
    """Row-sharded gridpp.optimal_interpolation over the (Y, X) grid given by `lats`, `lons` (and optional `elevs`,
    `lafs`): this rank analyses rows row_block(Y, world, rank) of `background` with the full observation set and
    returns them (the whole field on every rank with gather=True). Every rank passes the same arguments; the caller
    selects the rank's device beforehand (gridpp_b200.set_device). `compute(lats, lons, elevs, lafs, background)`, all
    restricted to the rank's rows, defaults to the CUDA path."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1 = row_block(len(lats), world, rank)
    args = [_rows_of(a, r0, r1) for a in (lats, lons, elevs, lafs, background)]
    if compute is None:
        import gridpp_b200 as gpp
        grid = gpp.Grid(args[0], args[1], args[2], args[3], type)
        rows = gpp.optimal_interpolation(grid, args[4], points, pobs, pratios, pbackground, structure, max_points, allow_extrapolation)
    else:
        rows = compute(*args)
    return _gather_result(rows, group) if gather else rows


def optimal_interpolation_ensi(lats, lons, background, points, pobs, psigmas, pbackground, structure, max_points,
                               allow_extrapolation=True, elevs=None, lafs=None, type=0, group=None, gather=False, compute=None):
    """Row-sharded gridpp.optimal_interpolation_ensi; `background` is (Y, X, E). A member with an invalid value anywhere
    in the WHOLE field is left untouched (oi_ensi.cpp:187-201), so the per-member flags of the ranks are combined with
    one all-reduce before the rows are analysed: members invalid elsewhere are masked in this rank's rows for the call
    and restored afterwards."""
    import numpy as np
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1 = row_block(len(lats), world, rank)
    args = [_rows_of(a, r0, r1) for a in (lats, lons, elevs, lafs)]
    mine = np.array(background[r0:r1], dtype=np.float32)
    invalid = (~np.isfinite(mine)).reshape(-1, mine.shape[-1]).any(axis=0)
    if world > 1:
        flags = torch.from_numpy(invalid.astype(np.int32))
        if dist.get_backend(group) == "nccl":
            flags = flags.cuda()
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
        invalid = flags.cpu().numpy().astype(bool)
    masked = mine.copy()
    masked[..., invalid] = np.nan         # marks the member invalid for this rank's call too
    if compute is None:
        import gridpp_b200 as gpp
        grid = gpp.Grid(args[0], args[1], args[2], args[3], type)
        rows = gpp.optimal_interpolation_ensi(grid, masked, points, pobs, psigmas, pbackground, structure, max_points, allow_extrapolation)
    else:
        rows = compute(*args, masked)
    rows = np.array(rows, dtype=np.float32)
    rows[..., invalid] = mine[..., invalid]
    return _gather_result(rows, group) if gather else rows
