"""Golden fixture for SURVEY.md 8f#1, generated from the UNMODIFIED reference sources (oracle/_ref/libgridpp_ref.so):

    make -C oracle ref && python tests/golden/make_golden_ensi_multi.py

gridpp::optimal_interpolation_ensi_multi_{ebe,ebesc,utem} (oi_ensi_multi.cpp) and gridpp::staticcorr_points (corr_points.cpp),
Points overloads, on a Cartesian and a Geodetic set-up with elevations and land-area fractions. The number of observations is
kept below the number of background points (the reference sizes its per-observation tables by the background, :422,:956).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def setup(rng, ctype, nB=400, nS=150, nE=8):
    if ctype == B.CARTESIAN:
        by, bx = rng.uniform(0, 40000, nB), rng.uniform(0, 40000, nB)
        py, px = rng.uniform(-1000, 41000, nS), rng.uniform(-1000, 41000, nS)
    else:
        by, bx = rng.uniform(59, 59.4, nB), rng.uniform(10, 10.8, nB)
        py, px = rng.uniform(58.99, 59.41, nS), rng.uniform(9.98, 10.82, nS)
    d = dict(by=by, bx=bx, py=py, px=px, be=rng.uniform(0, 400, nB), bf=rng.uniform(0, 1, nB), pe=rng.uniform(0, 400, nS),
             pf=rng.uniform(0, 1, nS))
    d = {k: v.astype(f32) for k, v in d.items()}
    trend = lambda y, x: 5 * np.sin(np.asarray(y, float) * (1 / 9000.0 if ctype == B.CARTESIAN else 40)) + 3 * np.cos(np.asarray(x, float) * (1 / 7000.0 if ctype == B.CARTESIAN else 25))
    d["background"] = (trend(by, bx)[:, None] + rng.normal(size=(nB, nE)) * 1.5).astype(f32)
    d["background_corr"] = (trend(by, bx)[:, None] * 0.5 + rng.normal(size=(nB, nE))).astype(f32)
    d["pbackground"] = (trend(py, px)[:, None] + rng.normal(size=(nS, nE)) * 1.5).astype(f32)
    d["pbackground_corr"] = (trend(py, px)[:, None] * 0.5 + rng.normal(size=(nS, nE))).astype(f32)
    d["pbackground_corr"][5] = 2.5          # no spread: the standardised perturbations are zeroed (:436-439)
    d["background_corr"][7] = -1.0
    truth = trend(py, px) + 1.0
    d["pobs2"] = (truth[:, None] + rng.normal(size=(nS, nE)) * 0.3).astype(f32)
    d["pobs2"][rng.uniform(size=nS) < 0.06, 0] = np.nan    # screened out by pobs[index][0] (:480)
    d["pobs1"] = d["pobs2"][:, 0].copy()
    d["pratios"] = rng.uniform(0.05, 0.6, nS).astype(f32)
    d["bratios"] = rng.uniform(0.7, 1.3, nB).astype(f32)
    return d


def main():
    ref = B.load("ref")
    ref.set_omp_threads(1)
    rng = np.random.default_rng(2024)
    store = {}
    for tag, ctype, spec in (("cart", B.CARTESIAN, (B.BARNES, 6000.0, 200.0, 0.5)), ("geo", B.GEODETIC, (B.CRESSMAN, 9000.0, 0.0, 0.0))):
        d = setup(rng, ctype)
        s = B.make_structure(*spec)
        store[tag + "__structure"] = np.array(spec, np.float64)
        store.update({tag + "__" + k: v for k, v in d.items()})
        bp, op = (d["by"], d["bx"], d["be"], d["bf"]), (d["py"], d["px"], d["pe"], d["pf"])
        for mp in (0, 12):
            store["%s__staticcorr_mp%d" % (tag, mp)] = ref.staticcorr_points(bp, op, s, mp, ctype)
            for extr in (0, 1):
                key = "%s__%%s_mp%d_x%d" % (tag, mp, extr)
                store[key % "ebesc"] = ref.ensi_multi("ebesc", bp, d["bratios"], d["background"], None, op, d["pobs2"], d["pratios"], d["pbackground"],
                                                      None, s, mp, ctype, bool(extr))
                store[key % "ebe"] = ref.ensi_multi("ebe", bp, d["bratios"], d["background"], d["background_corr"], op, d["pobs2"], d["pratios"],
                                                    d["pbackground"], d["pbackground_corr"], s, mp, ctype, bool(extr))
                store[key % "utem"] = ref.ensi_multi("utem", bp, d["bratios"], d["background"], d["background_corr"], op, d["pobs1"], d["pratios"],
                                                     d["pbackground"], d["pbackground_corr"], s, mp, ctype, bool(extr))
        # the last two members invalid somewhere: they are left untouched (the reference only supports a tail)
        bg = d["background"].copy()
        bg[3, 6] = np.nan
        pbg = d["pbackground"].copy()
        pbg[9, 7] = np.nan
        store[tag + "__tail_background"], store[tag + "__tail_pbackground"] = bg, pbg
        store[tag + "__tail_ebesc"] = ref.ensi_multi("ebesc", bp, d["bratios"], bg, None, op, d["pobs2"], d["pratios"], pbg, None, s, 12, ctype, False)
        store[tag + "__tail_ebe"] = ref.ensi_multi("ebe", bp, d["bratios"], bg, d["background_corr"], op, d["pobs2"], d["pratios"], pbg,
                                                   d["pbackground_corr"], s, 12, ctype, False)
        store[tag + "__tail_utem"] = ref.ensi_multi("utem", bp, d["bratios"], bg, d["background_corr"], op, d["pobs1"], d["pratios"], pbg,
                                                    d["pbackground_corr"], s, 12, ctype, False)
    np.savez_compressed(os.path.join(HERE, "ensi_multi.npz"), **store)
    print("wrote ensi_multi.npz (%d arrays)" % len(store))


if __name__ == "__main__":
    main()
