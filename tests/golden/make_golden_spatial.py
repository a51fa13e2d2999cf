"""Golden fixture for optimal_interpolation with SPATIALLY VARYING structure functions, generated from the UNMODIFIED
reference sources (oracle/_ref/libgridpp_ref.so):

    make -C oracle ref && python tests/golden/make_golden_spatial.py

<Family>Structure(Grid, vec2 h, vec2 v, vec2 w, min_rho): structure.cpp:168-184 (Barnes), :342 (Soar), :492 (Toar),
:643 (Powerlaw), :790 (Linear); used through gridpp::optimal_interpolation_full (oi.cpp:138-341).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)
    rng = np.random.default_rng(1000)
    f32 = np.float32
    ny, nx, dx = 36, 40, 1500.0                                  # background grid (Cartesian)
    yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    # elevations within 150 m: Soar / Toar use the SIGNED elevation difference (structure.cpp:52-53,62-63), so large negative
    # differences give "correlations" of order e^|dz/v| and nearly singular systems on which no two solvers agree to 1e-5
    belev = (rng.uniform(0, 150, size=(ny, nx))).astype(f32)
    blaf = rng.uniform(0, 1, size=(ny, nx)).astype(f32)
    bg = (rng.normal(size=(ny, nx)) * 2).astype(f32)
    S = 120
    py, px = rng.uniform(-3000, ny * dx + 3000, S).astype(f32), rng.uniform(-3000, nx * dx + 3000, S).astype(f32)
    pelev, plaf = rng.uniform(0, 150, S).astype(f32), rng.uniform(0, 1, S).astype(f32)
    pbg = rng.normal(size=S).astype(f32)
    obs = (pbg + rng.normal(size=S)).astype(f32)
    obs[7] = np.nan
    ratios = rng.uniform(0.2, 1.0, S).astype(f32)
    # scale grid: coarser than the background grid, and not aligned with it
    gny, gnx = 7, 9
    gy, gx = np.meshgrid(np.linspace(-2000, ny * dx + 2000, gny), np.linspace(-2000, nx * dx + 2000, gnx), indexing="ij")
    h = rng.uniform(3000, 9000, size=(gny, gnx)).astype(f32)
    v = rng.uniform(100, 400, size=(gny, gnx)).astype(f32)
    w = rng.uniform(0.3, 0.9, size=(gny, gnx)).astype(f32)
    v[2, 3] = 0          # a disabled vertical scale
    store = dict(y=yy.astype(f32), x=xx.astype(f32), belev=belev, blaf=blaf, background=bg, py=py, px=px, pelev=pelev, plaf=plaf,
                 pobs=obs, pratios=ratios, pbackground=pbg, gy=gy.astype(f32), gx=gx.astype(f32), h=h, v=v, w=w)
    cases = []
    for name, stype in (("barnes", B.BARNES), ("soar", B.SOAR), ("toar", B.TOAR), ("powerlaw", B.POWERLAW), ("linear", B.LINEAR)):
        for elev in (False, True):
            for mp, extr, min_rho in ((12, True, 0.0013), (0, False, 0.05)):
                if name == "linear" and not (elev and mp == 12):
                    continue                         # localization distance 0: one case is enough
                key = "%s__elev%d__mp%d" % (name, int(elev), mp)
                bpts = (yy, xx, belev if elev else None, blaf if elev else None)
                opts = (py, px, pelev if elev else None, plaf if elev else None)
                out, var = ref.optimal_interpolation_spatial(bpts, bg, opts, obs, ratios, pbg, stype, (gy, gx), h, v, w, min_rho, mp,
                                                             B.CARTESIAN, allow_extrapolation=extr, want_variance=True)
                store[key + "__analysis"] = out.reshape(ny, nx)
                store[key + "__variance"] = var.reshape(ny, nx)
                cases.append("%s,%d,%d,%d,%d,%g" % (name, stype, int(elev), mp, int(extr), min_rho))
    store["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "oi_spatial_structure.npz"), **store)
    print("wrote oi_spatial_structure.npz (%d cases)" % len(cases))
    changed = sum(float(np.abs(store[c.split(",")[0] + "__elev%s__mp%s__analysis" % (c.split(",")[2], c.split(",")[3])] - bg).max() > 0) for c in cases)
    print("cases whose analysis differs from the background:", int(changed), "of", len(cases))


if __name__ == "__main__":
    main()
