"""ctypes bindings for the two CPU checkers in this directory.

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg, never from the product package (gridpp_b200/).

  * ``load("ref")``    -> oracle/_ref/libgridpp_ref.so     the unmodified reference sources (prefix ``ref_``)
  * ``load("oracle")`` -> oracle/_ref/libgridpp_oracle.so  the plain-C restatement (prefix ``orc_``)

Both export the same flat functions (see oracle/ref_capi.cpp for the reference-side definitions and the
reference file:line each one wraps), so one binder serves both.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

GEODETIC, CARTESIAN = 0, 1
MEAN, MIN, MAX, SUM, COUNT = 0, 10, 30, 70, 80
MEDIAN, QUANTILE, STD, VARIANCE, RANDOMCHOICE = 20, 40, 50, 60, 90
BARNES, CRESSMAN, SOAR, TOAR, POWERLAW, LINEAR = range(6)


class StructureTerm(C.Structure):
    _fields_ = [("type", C.c_int), ("h", C.c_float), ("v", C.c_float), ("w", C.c_float),
                ("min_rho", C.c_float), ("loc_dist", C.c_float)]


class Structure(C.Structure):
    """Mirror of gpp_structure (include/gridpp_b200.h)."""
    _fields_ = [("n_terms", C.c_int), ("term", StructureTerm * 3), ("has_cv", C.c_int), ("cv_dist", C.c_float)]


_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def _f(a, shape=None):
    """float32 C-contiguous copy (or None)."""
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _p(a):
    if a is None:
        return None
    if a.dtype == np.float32:
        return a.ctypes.data_as(_fp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_ip)
    raise TypeError(a.dtype)


class CpuLib:
    def __init__(self, path, prefix):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self._fn("last_error").restype = C.c_char_p
        self._fn("calc_distance").restype = C.c_float
        self._fn("calc_distance").argtypes = [C.c_float] * 4 + [C.c_int]

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            msg = self._fn("last_error")().decode()
            raise (ValueError if rc == 1 else RuntimeError)(msg)

    # ---- threads
    def set_omp_threads(self, n):
        self._fn("set_omp_threads")(C.c_int(n))

    def get_omp_threads(self):
        return self._fn("get_omp_threads")()

    # ---- coordinates / structure
    def convert_coordinates(self, lats, lons, ctype):
        lats, lons = _f(lats).ravel(), _f(lons).ravel()
        n = lats.size
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        self._check(self._fn("convert_coordinates")(_p(lats), _p(lons), n, ctype, _p(x), _p(y), _p(z)))
        return x, y, z

    def structure_describe(self, stype, h, v=0.0, w=0.0, hmax=float("nan")):
        """localization distance of <Type>Structure(h, v, w, hmax) for a point"""
        out = C.c_float()
        self._check(self._fn("structure_describe")(stype, C.c_float(h), C.c_float(v), C.c_float(w), C.c_float(hmax),
                                                   C.byref(out)))
        return out.value

    def structure_corr(self, s, p1, p2, background=False):
        p1, p2 = _f(p1).reshape(-1, 5), _f(p2).reshape(-1, 5)
        n = p1.shape[0]
        out = np.empty(n, np.float32)
        self._check(self._fn("structure_corr")(C.byref(s), _p(p1), _p(p2), n, int(background), _p(out)))
        return out

    def structure_localization_distance(self, s):
        out = C.c_float()
        self._check(self._fn("structure_localization_distance")(C.byref(s), C.byref(out)))
        return out.value

    # ---- index queries
    def points_nearest(self, lats, lons, ctype, qlats, qlons, include_match=True, timing=None):
        lats, lons, qlats, qlons = _f(lats).ravel(), _f(lons).ravel(), _f(qlats).ravel(), _f(qlons).ravel()
        out = np.empty(qlats.size, np.int32)
        sec = C.c_double()
        self._check(self._fn("points_nearest")(_p(lats), _p(lons), lats.size, ctype, _p(qlats), _p(qlons), qlats.size,
                                               int(include_match), _p(out), C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out

    def points_neighbours(self, lats, lons, ctype, qlats, qlons, radii, include_match=True, capacity=64):
        lats, lons, qlats, qlons = _f(lats).ravel(), _f(lons).ravel(), _f(qlats).ravel(), _f(qlons).ravel()
        nq = qlats.size
        radii = _f(np.broadcast_to(np.asarray(radii, np.float32), (nq,)))
        idx = np.full((nq, capacity), -1, np.int32)
        dist = np.full((nq, capacity), np.nan, np.float32)
        cnt = np.zeros(nq, np.int32)
        self._check(self._fn("points_neighbours")(_p(lats), _p(lons), lats.size, ctype, _p(qlats), _p(qlons), _p(radii),
                                                  nq, int(include_match), capacity, _p(idx), _p(dist), _p(cnt)))
        return idx, dist, cnt

    def points_neighbours_raw(self, lats, lons, ctype, qlat, qlon, radius, include_match=True, capacity=64):
        lats, lons = _f(lats).ravel(), _f(lons).ravel()
        idx = np.full(capacity, -1, np.int32)
        cnt = C.c_int()
        self._check(self._fn("points_neighbours_raw")(_p(lats), _p(lons), lats.size, ctype, C.c_float(qlat),
                                                      C.c_float(qlon), C.c_float(radius), int(include_match), capacity,
                                                      _p(idx), C.byref(cnt)))
        return idx[:min(cnt.value, capacity)]

    def points_closest(self, lats, lons, ctype, qlats, qlons, num, include_match=True):
        lats, lons, qlats, qlons = _f(lats).ravel(), _f(lons).ravel(), _f(qlats).ravel(), _f(qlons).ravel()
        out = np.empty((qlats.size, num), np.int32)
        self._check(self._fn("points_closest")(_p(lats), _p(lons), lats.size, ctype, _p(qlats), _p(qlons), qlats.size, num,
                                               int(include_match), _p(out)))
        return out

    def calc_distance(self, lat1, lon1, lat2, lon2, ctype):
        return self._fn("calc_distance")(lat1, lon1, lat2, lon2, ctype)

    def nearest(self, ilats, ilons, ctype, qlats, qlons, ivalues, timing=None):
        ilats, ilons, qlats, qlons = _f(ilats).ravel(), _f(ilons).ravel(), _f(qlats).ravel(), _f(qlons).ravel()
        iv = _f(ivalues)
        single = iv.ndim == 1
        iv2 = iv.reshape(1, -1) if single else iv
        assert iv2.shape[1] == ilats.size
        out = np.empty((iv2.shape[0], qlats.size), np.float32)
        sec = C.c_double()
        self._check(self._fn("nearest")(_p(ilats), _p(ilons), ilats.size, ctype, _p(qlats), _p(qlons), qlats.size,
                                        _p(iv2), iv2.shape[0], _p(out), C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out[0] if single else out

    # ---- OI
    def optimal_interpolation(self, bpts, background, opts, pobs, obs_variance, pbackground, structure, max_points,
                              ctype, bvariance=None, bvariance_at_points=None, allow_extrapolation=True,
                              want_variance=False, timing=None):
        """bpts / opts: tuples (lats, lons, elevs|None, lafs|None), flattened."""
        bl, bo, be, bf = (_f(a).ravel() if a is not None else None for a in bpts)
        pl, po, pe, pf = (_f(a).ravel() if a is not None else None for a in opts)
        nB, nS = bl.size, pl.size
        bg = _f(background).ravel()
        bvar = _f(bvariance).ravel() if bvariance is not None else None
        obs, ovar, pbg = _f(pobs).ravel(), _f(obs_variance).ravel(), _f(pbackground).ravel()
        pbvar = _f(bvariance_at_points).ravel() if bvariance_at_points is not None else None
        assert bg.size == nB and obs.size == nS and ovar.size == nS and pbg.size == nS
        out = np.empty(nB, np.float32)
        var = np.empty(nB, np.float32)
        sec = C.c_double()
        self._check(self._fn("optimal_interpolation")(
            _p(bl), _p(bo), _p(be), _p(bf), nB, _p(bg), _p(bvar), _p(pl), _p(po), _p(pe), _p(pf), nS, ctype, _p(obs),
            _p(ovar), _p(pbg), _p(pbvar), C.byref(structure), max_points, int(allow_extrapolation), _p(out), _p(var),
            C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return (out, var) if want_variance else out

    def optimal_interpolation_grid(self, lats, lons, background, opts, pobs, pratios, pbackground, structure, max_points, ctype,
                                   allow_extrapolation=True, timing=None):
        """gridpp::optimal_interpolation(Grid...) oi.cpp:26-87 (compiled reference only): lats / lons / background are (Y, X)."""
        la, lo, bg = _f(lats), _f(lons), _f(background)
        ny, nx = la.shape
        pl, po = _f(opts[0]).ravel(), _f(opts[1]).ravel()
        nS = pl.size
        obs, rat, pbg = _f(pobs).ravel(), _f(pratios).ravel(), _f(pbackground).ravel()
        out = np.empty((ny, nx), np.float32)
        sec = C.c_double()
        self._check(self._fn("optimal_interpolation_grid")(_p(la), _p(lo), ny, nx, _p(bg), _p(pl), _p(po), nS, ctype, _p(obs), _p(rat),
                                                           _p(pbg), C.byref(structure), max_points, int(allow_extrapolation), _p(out),
                                                           C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out

    def optimal_interpolation_ensi(self, bpts, background, opts, pobs, psigmas, pbackground, structure, max_points, ctype,
                                   allow_extrapolation=True, timing=None):
        bl, bo, be, bf = (_f(a).ravel() if a is not None else None for a in bpts)
        pl, po, pe, pf = (_f(a).ravel() if a is not None else None for a in opts)
        nB, nS = bl.size, pl.size
        bg = _f(background).reshape(nB, -1)
        nE = bg.shape[1]
        pbg = _f(pbackground).reshape(nS, nE)
        obs, sig = _f(pobs).ravel(), _f(psigmas).ravel()
        out = np.empty((nB, nE), np.float32)
        sec = C.c_double()
        self._check(self._fn("optimal_interpolation_ensi")(
            _p(bl), _p(bo), _p(be), _p(bf), nB, _p(bg), nE, _p(pl), _p(po), _p(pe), _p(pf), nS, ctype, _p(obs), _p(sig),
            _p(pbg), C.byref(structure), max_points, int(allow_extrapolation), _p(out), C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out

    def staticcorr_points(self, pts, knots, structure, max_points, ctype):
        """gridpp::staticcorr_points (corr_points.cpp:26-131) -> (L, K)"""
        bl, bo, be, bf = (_f(a).ravel() if a is not None else None for a in pts)
        pl, po, pe, pf = (_f(a).ravel() if a is not None else None for a in knots)
        out = np.empty((bl.size, pl.size), np.float32)
        self._check(self._fn("staticcorr_points")(_p(bl), _p(bo), _p(be), _p(bf), bl.size, _p(pl), _p(po), _p(pe), _p(pf), pl.size, ctype,
                                                  C.byref(structure), max_points, _p(out)))
        return out

    def ensi_multi(self, kind, bpts, bratios, background, background_corr, opts, pobs, pratios, pbackground, pbackground_corr,
                   structure, max_points, ctype, allow_extrapolation=True):
        """optimal_interpolation_ensi_multi_{ebe,ebesc,utem} (Points overloads, oi_ensi_multi.cpp:329-1311); kind names the variant."""
        bl, bo, be, bf = (_f(a).ravel() if a is not None else None for a in bpts)
        pl, po, pe, pf = (_f(a).ravel() if a is not None else None for a in opts)
        nB, nS = bl.size, pl.size
        bg = _f(background).reshape(nB, -1)
        nE = bg.shape[1]
        pbg = _f(pbackground).reshape(nS, nE)
        bgc = _f(background_corr).reshape(nB, nE) if background_corr is not None else None
        pbgc = _f(pbackground_corr).reshape(nS, nE) if pbackground_corr is not None else None
        br, pr = _f(bratios).ravel(), _f(pratios).ravel()
        out = np.empty((nB, nE), np.float32)
        if kind == "utem":
            obs = _f(pobs).ravel()
            self._check(self._fn("ensi_multi_utem")(
                _p(bl), _p(bo), _p(be), _p(bf), nB, _p(br), _p(bg), _p(bgc), nE, _p(pl), _p(po), _p(pe), _p(pf), nS, ctype, _p(obs), _p(pr),
                _p(pbg), _p(pbgc), C.byref(structure), max_points, int(allow_extrapolation), _p(out)))
        else:
            obs = _f(pobs).reshape(nS, nE)
            self._check(self._fn("ensi_multi_ebe")(
                _p(bl), _p(bo), _p(be), _p(bf), nB, _p(br), _p(bg), _p(bgc), nE, _p(pl), _p(po), _p(pe), _p(pf), nS, ctype, _p(obs), _p(pr),
                _p(pbg), _p(pbgc), C.byref(structure), max_points, int(allow_extrapolation), int(kind == "ebe"), _p(out)))
        return out

    def optimal_interpolation_spatial(self, bpts, background, opts, pobs, pratios, pbackground, stype, sgrid, h, v, w, min_rho,
                                      max_points, ctype, allow_extrapolation=True, want_variance=False):
        """optimal_interpolation_full with <Family>Structure(Grid, h, v, w, min_rho); both checkers export it."""
        bl, bo, be, bf = (_f(a).ravel() if a is not None else None for a in bpts)
        pl, po, pe, pf = (_f(a).ravel() if a is not None else None for a in opts)
        nB, nS = bl.size, pl.size
        glats, glons = _f(sgrid[0]), _f(sgrid[1])
        gny, gnx = glats.shape
        out = np.empty(nB, np.float32)
        var = np.empty(nB, np.float32)
        self._check(self._fn("optimal_interpolation_spatial")(
            _p(bl), _p(bo), _p(be), _p(bf), nB, _p(_f(background).ravel()), _p(pl), _p(po), _p(pe), _p(pf), nS, ctype, _p(_f(pobs).ravel()),
            _p(_f(pratios).ravel()), _p(_f(pbackground).ravel()), int(stype), _p(glats), _p(glons), gny, gnx, _p(_f(h, (gny, gnx))),
            _p(_f(v, (gny, gnx))), _p(_f(w, (gny, gnx))), C.c_float(min_rho), max_points, int(allow_extrapolation), _p(out), _p(var)))
        return (out, var) if want_variance else out

    # ---- neighbourhood
    def neighbourhood(self, field, halfwidth, statistic, timing=None):
        f = _f(field)
        ny, nx = f.shape
        out = np.full((ny, nx), np.nan, np.float32)
        sec = C.c_double()
        self._check(self._fn("neighbourhood")(_p(f), ny, nx, halfwidth, statistic, _p(out), C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out

    def neighbourhood_search(self, array, search_array, halfwidth, tmin, tmax, delta, apply_array=None):
        """gridpp::neighbourhood_search (neighbourhood_search.cpp:7-113)"""
        a, s = _f(array), _f(search_array)
        ny, nx = a.shape
        ap = np.ascontiguousarray(apply_array, np.int32) if apply_array is not None else None
        out = np.empty((ny, nx), np.float32)
        self._check(self._fn("neighbourhood_search")(_p(a), _p(s), ny, nx, halfwidth, C.c_float(tmin), C.c_float(tmax), C.c_float(delta), _p(ap), _p(out)))
        return out

    def calc_gradient(self, base, values, gradient_type, halfwidth, num_min=2, min_range=float("nan"), default_gradient=0.0):
        """gridpp::calc_gradient (calc_gradient.cpp:6-126); gradient_type 0 = MinMax, 10 = LinearRegression"""
        b, v = _f(base), _f(values)
        ny, nx = b.shape
        out = np.empty((ny, nx), np.float32)
        self._check(self._fn("calc_gradient")(_p(b), _p(v), ny, nx, gradient_type, halfwidth, num_min, C.c_float(min_range), C.c_float(default_gradient), _p(out)))
        return out

    def neighbourhood_brute_force(self, field, halfwidth, statistic):
        f = _f(field)
        ny, nx = f.shape
        out = np.full((ny, nx), np.nan, np.float32)
        self._check(self._fn("neighbourhood_brute_force")(_p(f), ny, nx, halfwidth, statistic, _p(out)))
        return out

    def neighbourhood_window(self, field, halfwidth, statistic, quantile=0.0):
        """neighbourhood_brute_force / neighbourhood_quantile (statistic == QUANTILE) of a (Y, X) or (Y, X, E) field."""
        f = _f(field)
        ny, nx = f.shape[:2]
        ne = f.shape[2] if f.ndim == 3 else 1
        out = np.full((ny, nx), np.nan, np.float32)
        fn = self._fn("neighbourhood_window_ens")
        fn.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp]
        self._check(fn(_p(f), ny, nx, ne, halfwidth, statistic, float(quantile), _p(out)))
        return out

    # ---- consumers of the point index. Location sets are (lats, lons[, elevs]) arrays: 2-D -> a Grid, 1-D -> Points. The plain-C
    # oracle takes them flattened; the compiled reference builds the Grid / Points object the overload wants.
    def _sets(self, *sets):
        out = []
        for la, lo in sets:
            la, lo = _f(la), _f(lo)
            ny, nx = (la.shape[0], la.shape[1]) if la.ndim == 2 else (la.size, 0)
            out.append((la, lo, ny, nx))
        return out

    def gridding(self, oset, iset, values, radius, min_num, statistic, ctype, nearest=False):
        (ola, olo, ony, onx), (ila, ilo, iny, _) = self._sets(oset, iset)
        v = _f(values).ravel()
        out = np.full(ola.shape, np.nan, np.float32)
        name = "gridding_nearest" if nearest else "gridding"
        if self.prefix == "ref_":
            args = [_p(ola), _p(olo), ony, onx, _p(ila), _p(ilo), iny, ctype, _p(v)] + ([] if nearest else [C.c_float(radius)]) + [min_num, statistic, _p(out)]
        else:
            args = [_p(ola), _p(olo), ola.size, _p(ila), _p(ilo), iny, ctype, _p(v)] + ([] if nearest else [C.c_float(radius)]) + [min_num, statistic, _p(out)]
        self._check(self._fn(name)(*args))
        return out

    def count(self, iset, oset, radius, ctype):
        (ila, ilo, iny, inx), (ola, olo, ony, onx) = self._sets(iset, oset)
        out = np.full(ola.shape, np.nan, np.float32)
        if self.prefix == "ref_":
            self._check(self._fn("count")(_p(ila), _p(ilo), iny, inx, _p(ola), _p(olo), ony, onx, ctype, C.c_float(radius), _p(out)))
        else:
            self._check(self._fn("count")(_p(ila), _p(ilo), ila.size, _p(ola), _p(olo), ola.size, ctype, C.c_float(radius), _p(out)))
        return out

    def distance(self, iset, oset, num, ctype):
        (ila, ilo, iny, inx), (ola, olo, ony, onx) = self._sets(iset, oset)
        out = np.full(ola.shape, np.nan, np.float32)
        if self.prefix == "ref_":
            self._check(self._fn("distance")(_p(ila), _p(ilo), iny, inx, _p(ola), _p(olo), ony, onx, ctype, num, _p(out)))
        else:
            query_first = int(onx == 0)    # distance.cpp:21,111 vs :52,83
            self._check(self._fn("distance")(_p(ila), _p(ilo), ila.size, _p(ola), _p(olo), ola.size, ctype, num, query_first, _p(out)))
        return out

    def fill(self, glats, glons, field, plats, plons, radii, value, outside, ctype):
        la, lo, f = _f(glats), _f(glons), _f(field)
        ny, nx = la.shape
        pl, po, r = _f(plats).ravel(), _f(plons).ravel(), _f(radii).ravel()
        out = np.full((ny, nx), np.nan, np.float32)
        if self.prefix == "ref_":
            self._check(self._fn("fill")(_p(la), _p(lo), ny, nx, _p(f), _p(pl), _p(po), pl.size, ctype, _p(r), C.c_float(value), int(outside), _p(out)))
        else:
            self._check(self._fn("fill")(_p(la), _p(lo), ny * nx, _p(f), _p(pl), _p(po), pl.size, ctype, _p(r), C.c_float(value), int(outside), _p(out)))
        return out

    def fill_missing(self, field):
        f = _f(field)
        out = np.full(f.shape, np.nan, np.float32)
        self._check(self._fn("fill_missing")(_p(f), f.shape[0], f.shape[1], _p(out)))
        return out

    def doping(self, glats, glons, gelevs, background, plats, plons, pelevs, obs, extent, max_elev_diff, ctype, square):
        la, lo, bg = _f(glats), _f(glons), _f(background)
        ny, nx = la.shape
        ge = _f(gelevs) if gelevs is not None else np.full((ny, nx), np.nan, np.float32)
        pl, po, o = _f(plats).ravel(), _f(plons).ravel(), _f(obs).ravel()
        pe = _f(pelevs).ravel() if pelevs is not None else np.full(pl.size, np.nan, np.float32)
        out = np.full((ny, nx), np.nan, np.float32)
        if square:
            ext = np.ascontiguousarray(np.asarray(extent, np.int32).ravel())
            ext_p = ext.ctypes.data_as(_ip)
        else:
            ext = _f(extent).ravel()
            ext_p = _p(ext)
        name = "doping_square" if square else "doping_circle"
        if self.prefix == "ref_" or square:
            self._check(self._fn(name)(_p(la), _p(lo), _p(ge), ny, nx, _p(bg), _p(pl), _p(po), _p(pe), pl.size, ctype, _p(o), ext_p,
                                       C.c_float(max_elev_diff), _p(out)))
        else:
            self._check(self._fn(name)(_p(la), _p(lo), _p(ge), ny * nx, _p(bg), _p(pl), _p(po), _p(pe), pl.size, ctype, _p(o), ext_p,
                                       C.c_float(max_elev_diff), _p(out)))
        return out

    def neighbourhood_quantile_fast(self, field, quantile, halfwidth, thresholds, timing=None):
        f = _f(field)
        ny, nx = f.shape
        thr = _f(thresholds).ravel()
        qf = None
        q = float("nan")
        if np.ndim(quantile) == 0:
            q = float(quantile)
        else:
            qf = _f(quantile, (ny, nx))
        out = np.full((ny, nx), np.nan, np.float32)
        sec = C.c_double()
        self._check(self._fn("neighbourhood_quantile_fast")(_p(f), ny, nx, C.c_float(q), _p(qf), halfwidth, _p(thr),
                                                            thr.size, _p(out), C.byref(sec)))
        if timing is not None:
            timing.append(sec.value)
        return out

    def neighbourhood_ens(self, field, halfwidth, statistic):
        """gridpp::neighbourhood(vec3, halfwidth, statistic); field is (ny, nx, ne)."""
        f = _f(field)
        ny, nx, ne = f.shape
        out = np.full((ny, nx), np.nan, np.float32)
        self._check(self._fn("neighbourhood_ens")(_p(f), ny, nx, ne, halfwidth, statistic, _p(out)))
        return out

    def neighbourhood_quantile_fast_ens(self, field, quantile, halfwidth, thresholds):
        """gridpp::neighbourhood_quantile_fast(vec3, quantile | vec2, halfwidth, thresholds); field is (ny, nx, ne)."""
        f = _f(field)
        ny, nx, ne = f.shape
        thr = _f(thresholds).ravel()
        qf = None
        q = float("nan")
        if np.ndim(quantile) == 0:
            q = float(quantile)
        else:
            qf = _f(quantile, (ny, nx))
        out = np.full((ny, nx), np.nan, np.float32)
        self._check(self._fn("neighbourhood_quantile_fast_ens")(_p(f), ny, nx, ne, C.c_float(q), _p(qf), halfwidth, _p(thr), thr.size,
                                                                _p(out)))
        return out

    def get_neighbourhood_thresholds(self, field, num):
        f = _f(field)
        ny, nx = f.shape
        out = np.empty(max(num, 1), np.float32)
        n = C.c_int()
        self._check(self._fn("get_neighbourhood_thresholds")(_p(f), ny, nx, num, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def interpolate(self, x, ix, iy):
        ix, iy = _f(ix).ravel(), _f(iy).ravel()
        out = C.c_float()
        self._check(self._fn("interpolate")(C.c_float(x), _p(ix), _p(iy), ix.size, C.byref(out)))
        return out.value

    def calc_statistic(self, a, statistic):
        a = _f(a).ravel()
        out = C.c_float()
        self._check(self._fn("calc_statistic")(_p(a), a.size, statistic, C.byref(out)))
        return out.value

    def calc_quantile(self, a, q):
        a = _f(a).ravel()
        out = C.c_float()
        self._check(self._fn("calc_quantile")(_p(a), a.size, C.c_float(q), C.byref(out)))
        return out.value


_PATHS = {"ref": ("libgridpp_ref.so", "ref_"), "oracle": ("libgridpp_oracle.so", "orc_")}
_cache = {}


def available(kind):
    return os.path.exists(os.path.join(HERE, "_ref", _PATHS[kind][0]))


def load(kind):
    """kind: "ref" (compiled reference) or "oracle" (C restatement)."""
    if kind not in _cache:
        fname, prefix = _PATHS[kind]
        path = os.path.join(HERE, "_ref", fname)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        _cache[kind] = CpuLib(path, prefix)
    return _cache[kind]


def make_structure(stype=BARNES, h=0.0, v=0.0, w=0.0, hmax=float("nan"), min_rho=None):
    """Single-term descriptor. min_rho follows the reference constructors (structure.cpp:143-167, :317-340,
    :467-490, :618-641, :765-788); loc_dist is left 0 -- both CPU checkers recompute it themselves."""
    s = Structure()
    s.n_terms = 1
    t = s.term[0]
    t.type, t.h, t.v, t.w = stype, h, v, w
    if min_rho is None:
        # evaluated by the C oracle with the reference's own float/double mix (numpy's float32 exp is not glibc's expf)
        out = C.c_float()
        lib = load("oracle")
        lib._check(lib._fn("structure_min_rho")(stype, C.c_float(h), C.c_float(v), C.c_float(w), C.c_float(hmax), C.byref(out)))
        min_rho = out.value
    t.min_rho = float(min_rho)
    t.loc_dist = 0.0
    s.has_cv = 0
    s.cv_dist = float("nan")
    return s


def multiple_structure(sh, sv, sw):
    s = Structure()
    s.n_terms = 3
    for i, src in enumerate((sh, sv, sw)):
        for name, _ in StructureTerm._fields_:
            setattr(s.term[i], name, getattr(src.term[0], name))
    s.has_cv = 0
    s.cv_dist = float("nan")
    return s


def cross_validation(base, dist):
    s = Structure()
    C.memmove(C.byref(s), C.byref(base), C.sizeof(Structure))
    s.has_cv = 1
    s.cv_dist = dist
    return s
