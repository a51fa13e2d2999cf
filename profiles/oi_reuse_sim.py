"""Offline model (numpy) of the solved-system reuse of oi_fast_kernel on the C3 geometry: for a sub-grid, the set of the
max_points observations with the largest correlation is computed for every grid point, and the number of systems a warp
has to assemble and eliminate is counted for a given chunk shape (in 4 x 4-point tiles, serpentine order inside the chunk)
and LRU size. Reproduces the measured rates (profiles/oi_stats.py on the B200: 1 x 16 tiles / LRU 4 -> 8.15 %, 4 x 4 / 4
-> 7.24 %, 4 x 4 / 8 -> 5.75 %) and predicts other shapes.  usage: python profiles/oi_reuse_sim.py"""
import numpy as np

rng = np.random.default_rng(1000)
n, dx, R, K = 512, 250.0, 36456.5, 30
ext = n * dx
S = int(0.01e-6 * (ext + 2 * R) ** 2)
px, py = rng.uniform(-R, ext + R, S), rng.uniform(-R, ext + R, S)
gy, gx = np.meshgrid(np.arange(n) * dx, np.arange(n) * dx, indexing="ij")
ids = np.empty(n * n, np.int64)
table = {}
for r0 in range(0, n, 16):
    yy = gy[r0:r0 + 16].ravel()[:, None]
    xx = gx[r0:r0 + 16].ravel()[:, None]
    d2 = (yy - py[None]) ** 2 + (xx - px[None]) ** 2
    idx = np.argpartition(d2, K, axis=1)[:, :K]
    idx = np.where(np.take_along_axis(d2, idx, axis=1) <= R * R, idx, -1)
    idx.sort(axis=1)
    for j, row in enumerate(idx):
        ids[r0 * n + j] = table.setdefault(row.tobytes(), len(table))
ids = ids.reshape(n, n)
print("%d x %d points, %d observations, %d distinct selections (%.2f %% of the points)" % (n, n, S, len(table), 100.0 * len(table) / ids.size))


def solves(cty, ctx, lru_size):
    tiles = n // 4
    count = 0
    for cy in range(0, tiles, cty):
        for cx in range(0, tiles, ctx):
            lru = []
            prev = -1
            for tyy in range(cty):
                xs = range(ctx) if tyy % 2 == 0 else range(ctx - 1, -1, -1)
                for txx in xs:
                    ty, tx = cy + tyy, cx + txx
                    if ty >= tiles or tx >= tiles:
                        continue
                    for i in range(4):
                        cols = range(4) if i % 2 == 0 else range(3, -1, -1)
                        for j in cols:
                            s = ids[4 * ty + i, 4 * tx + j]
                            if s == prev:
                                continue
                            prev = s
                            if s in lru:
                                lru.remove(s)
                            else:
                                count += 1
                                if len(lru) == lru_size:
                                    lru.pop(0)
                            lru.append(s)
    return 100.0 * count / ids.size


for cty, ctx, lru in ((1, 16, 4), (4, 4, 4), (4, 4, 8), (4, 4, 16), (8, 8, 8), (8, 8, 16), (16, 16, 16), (8, 8, 32)):
    print("chunk %2d x %2d tiles, LRU %2d: %.2f %% of the points solve a system" % (cty, ctx, lru, solves(cty, ctx, lru)))


def hilbert_order(order):
    """(y, x) cells of a 2^order square along the Hilbert curve."""
    pts = []
    size = 1 << order
    for d in range(size * size):
        x = y = 0
        t = d
        s = 1
        while s < size:
            rx = 1 & (t // 2)
            ry = 1 & (t ^ rx)
            if ry == 0:
                if rx == 1:
                    x, y = s - 1 - x, s - 1 - y
                x, y = y, x
            x += s * rx
            y += s * ry
            t //= 4
            s *= 2
        pts.append((y, x))
    return pts


def solves_curve(order, lru_size, point_level):
    """Chunks of 2^order x 2^order tiles walked along a Hilbert curve (over tiles, or over the points themselves)."""
    side = (4 << order) if point_level else (1 << order)
    curve = hilbert_order(order + 2 if point_level else order)
    count = 0
    step = side if point_level else side * 4
    for y0 in range(0, n, step):
        for x0 in range(0, n, step):
            lru = []
            prev = -1
            cells = [(y0 + y, x0 + x) for y, x in curve] if point_level else \
                    [(y0 + 4 * ty + i, x0 + 4 * tx + (j if i % 2 == 0 else 3 - j)) for ty, tx in curve for i in range(4) for j in range(4)]
            for y, x in cells:
                if y >= n or x >= n:
                    continue
                s = ids[y, x]
                if s == prev:
                    continue
                prev = s
                if s in lru:
                    lru.remove(s)
                else:
                    count += 1
                    if len(lru) == lru_size:
                        lru.pop(0)
                lru.append(s)
    return 100.0 * count / ids.size


for order, lru in ((2, 8), (3, 8), (3, 16)):
    print("chunk %2d x %2d tiles along a Hilbert curve over tiles, LRU %2d: %.2f %%" % (1 << order, 1 << order, lru, solves_curve(order, lru, False)))
