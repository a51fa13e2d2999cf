"""gridpp_b200 -- B200-native drop-in for the data-parallel hot path of metno/gridpp.

The module mirrors the Python surface the reference generates with SWIG (swig/gridpp.i) for that path: same
names, argument order, defaults, return types (new float32 / int32 ndarrays) and exception types
(``ValueError`` for invalid arguments, ``RuntimeError`` otherwise)::

    import gridpp_b200 as gridpp
    grid = gridpp.Grid(lats, lons)
    points = gridpp.Points(plats, plons)
    structure = gridpp.BarnesStructure(10000)
    pbackground = gridpp.nearest(grid, points, background)
    analysis = gridpp.optimal_interpolation(grid, background, points, obs, ratios, pbackground, structure, 30)
    smooth = gridpp.neighbourhood(analysis, 7, gridpp.Mean)

Every function is a thin marshalling layer over the C ABI of ``libgridpp_b200.so`` (include/gridpp_b200.h);
all computation happens in hand-written sm_100a CUDA kernels. There is no CPU fallback.
"""
import ctypes as _C
import math as _math

import numpy as _np

from . import _lib
from ._lib import NotImplementedOnDevice, check as _check, lib as _libc

__version__ = _libc.gpp_version().decode()

# gridpp::CoordinateType (gridpp.h:120-123)
Geodetic = 0
Cartesian = 1
# gridpp::Statistic (gridpp.h:88-100)
Mean, Min, Median, Max, Quantile, Std, Variance, Sum, Count, RandomChoice, Unknown = 0, 10, 20, 30, 40, 50, 60, 70, 80, 90, -1
MV = float("nan")
# gridpp::GradientType (gridpp.h:126-129)
MinMax, LinearRegression = 0, 10
_BARNES, _CRESSMAN, _SOAR, _TOAR, _POWERLAW, _LINEAR = range(6)


def version():
    return __version__


def device_count():
    n = _C.c_int()
    _check(_libc.gpp_device_count(_C.byref(n)))
    return n.value


def set_device(device):
    _check(_libc.gpp_set_device(int(device)))


def synchronize():
    _check(_libc.gpp_device_synchronize())


def kernel_launch_count():
    return int(_libc.gpp_kernel_launch_count())


def measure_fp64_fma_peak():
    """Measured fp64 FMA throughput of the current device, TFLOP/s."""
    t = _C.c_double()
    _check(_libc.gpp_measure_fp64_fma_peak(_C.byref(t)))
    return t.value


# ---------------------------------------------------------------------------------------------------------
# marshalling helpers (the SWIG typemaps of swig/vector.i: any array-like of any dtype comes in, float32 /
# int32 C-contiguous arrays go out; a wrong number of dimensions is an error)
def _farray(a, ndim, name):
    try:
        arr = _np.ascontiguousarray(a, dtype=_np.float32)
    except (TypeError, ValueError) as e:
        raise ValueError("%s: cannot convert to a float array (%s)" % (name, e))
    if arr.ndim != ndim:
        # zero-length inputs such as [] or [[]] are accepted in any rank (tests/test_swig.py:91-107)
        if arr.size == 0:
            return arr.reshape((0,) * ndim)
        raise ValueError("%s must have %d dimension(s), got %d" % (name, ndim, arr.ndim))
    return arr


def _fptr(a):
    return a.ctypes.data_as(_lib.fp) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(_lib.ip) if a is not None else None


# ---------------------------------------------------------------------------------------------------------
class _PointSet:
    """Owner of a gpp_points handle."""

    def __init__(self, lats, lons, elevs, lafs, ctype):
        n = lats.size
        self._lats, self._lons = lats, lons
        self._elevs = elevs if elevs is not None else _np.full(n, _np.nan, _np.float32)
        self._lafs = lafs if lafs is not None else _np.full(n, _np.nan, _np.float32)
        self._type = int(ctype)
        self._handle = _C.c_void_p()
        _check(_libc.gpp_points_create(_fptr(lats), _fptr(lons), _fptr(elevs), _fptr(lafs), n, self._type,
                                       _C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h and _libc is not None:   # _libc is None during interpreter shutdown
            _libc.gpp_points_destroy(h)
            self._handle = None

    @property
    def n(self):
        return self._lats.size

    def xyz(self):
        x, y, z = (_np.empty(self.n, _np.float32) for _ in range(3))
        _check(_libc.gpp_points_get_xyz(self._handle, _fptr(x), _fptr(y), _fptr(z)))
        return x, y, z

    def nearest(self, lats, lons, include_match=True):
        lats, lons = _farray(lats, 1, "lats"), _farray(lons, 1, "lons")
        out = _np.empty(lats.size, _np.int32)
        _check(_libc.gpp_points_nearest_host(self._handle, _fptr(lats), _fptr(lons), lats.size, int(include_match), _iptr(out)))
        return out

    def closest(self, lats, lons, num, include_match=True):
        lats, lons = _farray(lats, 1, "lats"), _farray(lons, 1, "lons")
        out = _np.empty((lats.size, num), _np.int32)
        _check(_libc.gpp_points_closest_host(self._handle, _fptr(lats), _fptr(lons), lats.size, int(num), int(include_match),
                                             _iptr(out)))
        return out

    def neighbours(self, lats, lons, radii, include_match=True, capacity=None, with_distance=False):
        lats, lons = _farray(lats, 1, "lats"), _farray(lons, 1, "lons")
        nq = lats.size
        radii = _np.ascontiguousarray(_np.broadcast_to(_np.asarray(radii, _np.float32), (nq,)))
        count = _np.zeros(nq, _np.int32)
        if capacity is None or capacity == 0:
            _check(_libc.gpp_points_neighbours_host(self._handle, _fptr(lats), _fptr(lons), _fptr(radii), nq,
                                                    int(include_match), 0, None, None, _iptr(count)))
            if capacity is None:
                capacity = int(count.max()) if nq else 0
        idx = _np.full((nq, capacity), -1, _np.int32)
        dist = _np.full((nq, capacity), _np.nan, _np.float32) if with_distance else None
        if capacity > 0:
            _check(_libc.gpp_points_neighbours_host(self._handle, _fptr(lats), _fptr(lons), _fptr(radii), nq,
                                                    int(include_match), capacity, _iptr(idx), _fptr(dist), _iptr(count)))
        return idx, dist, count


class Points:
    """gridpp::Points (gridpp.h:1876-1968, points.cpp)."""

    def __init__(self, lats=None, lons=None, elevs=None, lafs=None, type=Geodetic):
        lats = _farray([] if lats is None else lats, 1, "lats")
        lons = _farray([] if lons is None else lons, 1, "lons")
        n = lats.size
        if lons.size != n:
            raise ValueError("Cannot create points with unequal lat and lon sizes")
        elevs = _farray([] if elevs is None else elevs, 1, "elevs")
        lafs = _farray([] if lafs is None else lafs, 1, "lafs")
        if elevs.size not in (0, n):
            raise ValueError("'elevs' must either be size 0 or the same size at lats/lons")
        if lafs.size not in (0, n):
            raise ValueError("'lafs' must either be size 0 or the same size at lats/lons")
        self._set = _PointSet(lats, lons, elevs if elevs.size == n and n > 0 else None, lafs if lafs.size == n and n > 0 else None, type)

    # -- accessors
    def get_lats(self):
        return self._set._lats.copy()

    def get_lons(self):
        return self._set._lons.copy()

    def get_elevs(self):
        return self._set._elevs.copy()

    def get_lafs(self):
        return self._set._lafs.copy()

    def size(self):
        return self._set.n

    def get_coordinate_type(self):
        return self._set._type

    # -- queries (points.cpp:40-62)
    def get_nearest_neighbour(self, lat, lon, include_match=True):
        return int(self._set.nearest([lat], [lon], include_match)[0])

    def get_closest_neighbours(self, lat, lon, num, include_match=True):
        row = self._set.closest([lat], [lon], num, include_match)[0]
        return row[row >= 0].copy()

    def get_neighbours(self, lat, lon, radius, include_match=True):
        idx, _, count = self._set.neighbours([lat], [lon], radius, include_match)
        return idx[0, :count[0]].copy()

    def get_neighbours_with_distance(self, lat, lon, radius, include_match=True):
        idx, dist, count = self._set.neighbours([lat], [lon], radius, include_match, with_distance=True)
        return idx[0, :count[0]].copy(), dist[0, :count[0]].copy()

    def get_num_neighbours(self, lat, lon, radius, include_match=True):
        _, _, count = self._set.neighbours([lat], [lon], radius, include_match, capacity=0)
        return int(count[0])

    def get_point(self, index):
        """points.cpp:128-130: the point with the coordinates of the index."""
        s = self._set
        if not 0 <= index < s.n:
            raise ValueError("Point index %d out of range" % index)
        _, x, y, z = convert_coordinates(s._lats[index], s._lons[index], s._type)
        return Point(s._lats[index], s._lons[index], s._elevs[index], s._lafs[index], s._type, x, y, z)

    def get_in_domain_indices(self, grid):
        """Points::get_in_domain_indices, points.cpp:77-92: the points that lie inside a cell of `grid` (Grid::get_box)."""
        found = grid._boxes(self._set._lats, self._set._lons)[0]
        return _np.nonzero(found)[0].astype(_np.int32)

    def get_in_domain(self, grid):
        """Points::get_in_domain, points.cpp:93-110 (the new set has the default coordinate type, like subset())."""
        return self.subset(self.get_in_domain_indices(grid))

    def subset(self, indices):
        indices = _np.asarray(indices, dtype=_np.int64).ravel()
        if indices.size and indices.max() >= self.size():
            raise ValueError("Index %d exceeds number of points %d" % (indices.max(), self.size()))
        s = self._set
        # points.cpp:132-150: the subset is created with the default (Geodetic) coordinate type
        return Points(s._lats[indices], s._lons[indices], s._elevs[indices], s._lafs[indices])


class Grid:
    """gridpp::Grid (gridpp.h:1971-2060, grid.cpp): a 2-D array of points, flattened row-major for queries."""

    def __init__(self, lats=None, lons=None, elevs=None, lafs=None, type=Geodetic):
        lats = _farray(_np.zeros((0, 0), _np.float32) if lats is None else lats, 2, "lats")
        lons = _farray(_np.zeros((0, 0), _np.float32) if lons is None else lons, 2, "lons")
        if lats.shape != lons.shape:
            raise ValueError("Cannot create grid with unequal lat and lon sizes")
        self._shape = lats.shape if lats.size else (0, 0)      # Grid::size(), grid.cpp:122-130: (0, 0) without nodes
        self._stored_shape = lats.shape                        # what get_lats() etc. return: the arrays as given (grid.cpp:12-55)
        elevs = _farray([[]] if elevs is None else elevs, 2, "elevs")
        lafs = _farray([[]] if lafs is None else lafs, 2, "lafs")
        # grid.cpp:41-54: elevations / land fractions of the wrong shape are replaced by missing values
        e = elevs.ravel() if elevs.shape == lats.shape and lats.size else None
        l = lafs.ravel() if lafs.shape == lats.shape and lats.size else None
        self._set = _PointSet(lats.ravel(), lons.ravel(), e, l, type)
        if lats.size:
            _check(_libc.gpp_points_set_shape(self._set._handle, self._shape[0], self._shape[1]))

    def size(self):
        return _np.array(self._shape, dtype=_np.int32)

    def get_lats(self):
        return self._set._lats.reshape(self._stored_shape).copy()

    def get_lons(self):
        return self._set._lons.reshape(self._stored_shape).copy()

    def get_elevs(self):
        return self._set._elevs.reshape(self._stored_shape).copy()

    def get_lafs(self):
        return self._set._lafs.reshape(self._stored_shape).copy()

    def get_coordinate_type(self):
        return self._set._type

    def _unflatten(self, flat):
        flat = _np.asarray(flat)
        nx = self._shape[1]
        return _np.stack([flat // nx, flat % nx], axis=-1).astype(_np.int32)   # grid.cpp:108-114

    def get_nearest_neighbour(self, lat, lon, include_match=True):
        i = int(self._set.nearest([lat], [lon], include_match)[0])
        return self._unflatten(i) if i >= 0 else _np.zeros(0, _np.int32)

    def get_closest_neighbours(self, lat, lon, num, include_match=True):
        row = self._set.closest([lat], [lon], num, include_match)[0]
        return self._unflatten(row[row >= 0]).reshape(-1, 2)

    def get_neighbours(self, lat, lon, radius, include_match=True):
        idx, _, count = self._set.neighbours([lat], [lon], radius, include_match)
        if count[0] == 0:
            return _np.zeros((0, 0), _np.int32)     # an empty ivec2
        return self._unflatten(idx[0, :count[0]]).reshape(-1, 2)

    def get_neighbours_with_distance(self, lat, lon, radius, include_match=True):
        idx, dist, count = self._set.neighbours([lat], [lon], radius, include_match, with_distance=True)
        return self._unflatten(idx[0, :count[0]]).reshape(-1, 2), dist[0, :count[0]].copy()

    def get_num_neighbours(self, lat, lon, radius, include_match=True):
        _, _, count = self._set.neighbours([lat], [lon], radius, include_match, capacity=0)
        return int(count[0])

    def get_point(self, y_index, x_index):
        """grid.cpp:232-235."""
        if not (0 <= y_index < self._shape[0] and 0 <= x_index < self._shape[1]):
            raise ValueError("Grid index (%d, %d) out of range" % (y_index, x_index))
        return Points.get_point(self, y_index * self._shape[1] + x_index)

    def to_points(self):
        p = Points.__new__(Points)
        p._set = self._set   # grid.cpp:131-145: same tree, elevations and land fractions
        return p

    def _boxes(self, lats, lons):
        """Grid::get_box (grid.cpp:149-231) for many locations: the grid cell (Y1, X1, Y2, X2) containing each one, found
        from its nearest node (one batched device query) and the reference's point_in_rectangle test
        (util.cpp:566-578, float arithmetic) on the up to four cells around that node. Returns (found, Y1, X1, Y2, X2)."""
        f = _np.float32
        lats, lons = _np.asarray(lats, f).ravel(), _np.asarray(lons, f).ravel()
        n = lats.size
        found = _np.zeros(n, bool)
        box = _np.full((4, n), -1, _np.int32)
        nY, nX = self._shape
        if n == 0 or self._set.n == 0 or nX <= 1 or nY <= 1:
            return (found,) + tuple(box)
        nn = self._set.nearest(lats, lons, True).astype(_np.int64)
        ok = nn >= 0
        Y, X = nn // nX, nn % nX
        la, lo = self._set._lats.reshape(self._shape), self._set._lons.reshape(self._shape)

        def side(a_lat, a_lon, b_lat, b_lon):   # D = (AB.lat * m.lon + AB.lon * m.lat) + C, AB = vect2d(A, B)
            ab_lon, ab_lat = f(b_lon - a_lon), f(f(-1) * f(b_lat - a_lat))
            c = f(f(-1) * f(f(ab_lat * a_lon) + f(ab_lon * a_lat)))
            return f(f(f(ab_lat * lons) + f(ab_lon * lats)) + c)

        for it in range(4):                     # the reference's order: (ydir, xdir) = (1,-1), (1,1), (-1,-1), (-1,1)
            xdir, ydir = -1 + 2 * (it % 2), -1 + 2 * (it < 2)
            cand = ok & ~found & ~((Y == 0) & (ydir == -1)) & ~((Y == nY - 1) & (ydir == 1)) & ~((X == 0) & (xdir == -1)) & ~((X == nX - 1) & (xdir == 1))
            if not cand.any():
                continue
            Yc, Xc = _np.clip(Y + ydir, 0, nY - 1), _np.clip(X + xdir, 0, nX - 1)
            A = (la[Y, X], lo[Y, X]); B = (la[Yc, X], lo[Yc, X]); C = (la[Yc, Xc], lo[Yc, Xc]); D = (la[Y, Xc], lo[Y, Xc])
            with _np.errstate(all="ignore"):
                d1, d2, d3, d4 = side(*A, *B), side(*A, *D), side(*B, *C), side(*C, *D)
            inside = ((0 >= d1) & (0 >= d4) & (0 <= d2) & (0 >= d3)) | ((0 <= d1) & (0 <= d4) & (0 >= d2) & (0 <= d3))
            hit = cand & inside
            box[0][hit] = _np.where(ydir == 1, Y, Y - 1)[hit]
            box[2][hit] = _np.where(ydir == 1, Y + 1, Y)[hit]
            box[1][hit] = _np.where(xdir == 1, X, X - 1)[hit]
            box[3][hit] = _np.where(xdir == 1, X + 1, X)[hit]
            found |= hit
        return (found,) + tuple(box)

    def get_box(self, lat, lon):
        """Grid::get_box: (found, Y1, X1, Y2, X2), as the SWIG wrapper returns the reference's output arguments."""
        found, y1, x1, y2, x2 = self._boxes([lat], [lon])
        return bool(found[0]), int(y1[0]), int(x1[0]), int(y2[0]), int(x2[0])


class KDTree(Points):
    """gridpp::KDTree (gridpp.h:1746-1873): the index without elevation / land-fraction metadata."""

    def __init__(self, lats=None, lons=None, type=Geodetic):
        Points.__init__(self, lats, lons, None, None, type)

    def get_x(self):
        return self._set.xyz()[0]

    def get_y(self):
        return self._set.xyz()[1]

    def get_z(self):
        return self._set.xyz()[2]

    # ---- scalar helpers of kdtree.cpp:107-200 (host arithmetic with the reference's float / double mix; SWIG also
    # exposes them flat as gridpp.KDTree_calc_distance etc., tests/test_kdtree.py:56,111)
    @staticmethod
    def deg2rad(deg):
        return float(_np.float32(float(_np.float32(deg)) * _math.pi / 180))

    @staticmethod
    def rad2deg(rad):
        return float(_np.float32(float(_np.float32(rad)) * 180 / _math.pi))

    @staticmethod
    def calc_straight_distance(x0, y0=None, z0=None, x1=None, y1=None, z1=None):
        if y0 is not None and z0 is None:                       # (Point, Point), kdtree.cpp:189-191
            p1, p2 = x0, y0
            x0, y0, z0, x1, y1, z1 = p1.x, p1.y, p1.z, p2.x, p2.y, p2.z
        f = _np.float32
        dx, dy, dz = f(x0) - f(x1), f(y0) - f(y1), f(z0) - f(z1)
        return float(_np.sqrt(dx * dx + dy * dy + dz * dz))

    @staticmethod
    def calc_distance(lat1, lon1, lat2=None, lon2=None, type=Geodetic):
        if lat2 is None:                                        # (Point, Point), kdtree.cpp:183-188
            p1, p2 = lat1, lon1
            if p1.type != p2.type:
                raise RuntimeError("Coordinate types must be the same")
            lat1, lon1, lat2, lon2, type = p1.lat, p1.lon, p2.lat, p2.lon, p1.type
        f = _np.float32
        lat1, lon1, lat2, lon2 = f(lat1), f(lon1), f(lat2), f(lon2)
        if type == Cartesian:
            dx, dy = lon1 - lon2, lat1 - lat2
            return float(_np.sqrt(dx * dx + dy * dy))
        if lat1 == lat2 and lon1 == lon2:
            return 0.0
        a1, a2, o1, o2 = KDTree.deg2rad(lat1), KDTree.deg2rad(lat2), KDTree.deg2rad(lon1), KDTree.deg2rad(lon2)
        cos, sin = _math.cos, _math.sin
        ratio = cos(a1) * cos(o1) * cos(a2) * cos(o2) + cos(a1) * sin(o1) * cos(a2) * sin(o2) + sin(a1) * sin(a2)
        try:
            dist = _math.acos(ratio) * 6.378137e6
        except ValueError:                                      # ratio a rounding above 1: acos gives NaN in C
            dist = float("nan")
        return float(f(dist))


def _calc_distance_fast(lat1, lon1, lat2=None, lon2=None, type=Geodetic):
    """KDTree::calc_distance_fast, kdtree.cpp:134-180 (equirectangular approximation; (Point, Point) form :181-186)."""
    if lat2 is None:
        p1, p2 = lat1, lon1
        if p1.type != p2.type:
            raise RuntimeError("Coordinate types must be the same")
        lat1, lon1, lat2, lon2, type = p1.lat, p1.lon, p2.lat, p2.lon, p1.type
    f = _np.float32
    lat1, lon1, lat2, lon2 = f(lat1), f(lon1), f(lat2), f(lon2)
    if type == Cartesian:
        dx, dy = lon1 - lon2, lat1 - lat2
        return float(_np.sqrt(dx * dx + dy * dy))
    lat1r, lat2r, lon1r, lon2r = KDTree.deg2rad(lat1), KDTree.deg2rad(lat2), KDTree.deg2rad(lon1), KDTree.deg2rad(lon2)
    dlon = _math.fmod(abs(lon1r - lon2r), 2 * _math.pi)
    if dlon > _math.pi:
        dlon = 2 * _math.pi - dlon
    max_lat = lat2r if abs(lat2r) > abs(lat1r) else lat1r
    dx2 = f(_math.cos(max_lat) ** 2 * dlon * dlon)
    dy2 = f((lat1r - lat2r) * (lat1r - lat2r))
    return float(f(6.378137e6 * _math.sqrt(float(f(dx2 + dy2)))))


KDTree.calc_distance_fast = staticmethod(_calc_distance_fast)
KDTree_calc_distance_fast = _calc_distance_fast
KDTree_deg2rad, KDTree_rad2deg = KDTree.deg2rad, KDTree.rad2deg
KDTree_calc_distance, KDTree_calc_straight_distance = KDTree.calc_distance, KDTree.calc_straight_distance


def is_valid(value):
    """gridpp::is_valid, util.cpp:16-18."""
    return not (_math.isnan(value) or _math.isinf(value))


# gridpp.cpp:45-68. The device path has no host thread team; the value is kept so that code which sets it and reads it
# back behaves as before.
_omp_threads = [1]


def set_omp_threads(num):
    _omp_threads[0] = int(num)


def get_omp_threads():
    return _omp_threads[0]


def initialize_omp():
    pass


# ---------------------------------------------------------------------------------------------------------
def convert_coordinates(lats, lons, type):
    """gridpp::convert_coordinates, util.cpp:583-615. As the SWIG module, returns (status, x, y, z): floats for scalar
    arguments (tests/test_points.py:130), arrays for array arguments."""
    scalar = _np.ndim(lats) == 0 and _np.ndim(lons) == 0
    la, lo = _farray(_np.atleast_1d(lats), 1, "lats"), _farray(_np.atleast_1d(lons), 1, "lons")
    if la.size != lo.size:
        raise ValueError("Cannot convert coordinates with unequal lat and lon sizes")
    x, y, z = (_np.empty(la.size, _np.float32) for _ in range(3))
    _check(_libc.gpp_convert_coordinates(_fptr(la), _fptr(lo), la.size, int(type), _fptr(x), _fptr(y), _fptr(z)))
    if scalar:
        return True, float(x[0]), float(y[0]), float(z[0])
    return True, x, y, z


class Point:
    """gridpp::Point (gridpp.h:1713-1743, point.cpp:5-26): Point(lat, lon, elev=MV, laf=MV, type=Geodetic) or, with the
    3-D coordinates given, Point(lat, lon, elev, laf, type, x, y, z)."""

    def __init__(self, lat, lon, elev=MV, laf=MV, type=Geodetic, x=None, y=None, z=None):
        self.lat, self.lon, self.elev, self.laf, self.type = float(lat), float(lon), float(elev), float(laf), int(type)
        if x is not None:
            self.x, self.y, self.z = float(x), float(y), float(z)
        elif self.type == Geodetic:
            xyz = [_np.empty(1, _np.float32) for _ in range(3)]
            la, lo = _np.array([lat], _np.float32), _np.array([lon], _np.float32)
            _check(_libc.gpp_convert_coordinates(_fptr(la), _fptr(lo), 1, Geodetic, _fptr(xyz[0]), _fptr(xyz[1]), _fptr(xyz[2])))
            self.x, self.y, self.z = (float(a[0]) for a in xyz)
        else:
            # point.cpp:18-22: x = lat, y = lon (not the x = lon, y = lat of convert_coordinates) and no validity check
            self.x, self.y, self.z = float(_np.float32(lat)), float(_np.float32(lon)), 0.0

    def _row(self):
        return [self.x, self.y, self.z, self.elev, self.laf]


class StructureFunction:
    """POD-backed replacement for the gridpp::StructureFunction hierarchy (gridpp.h:2069-2343)."""

    def __init__(self, desc):
        self._desc = desc

    @staticmethod
    def _points5(p):
        if isinstance(p, Point):
            p = [p._row()]
        elif isinstance(p, (list, tuple)) and len(p) > 0 and isinstance(p[0], Point):
            p = [q._row() for q in p]
        return _np.ascontiguousarray(_np.asarray(p, _np.float32).reshape(-1, 5))

    def corr(self, p1, p2):
        """corr(Point, Point) -> float and corr(Point, [Point, ...]) -> array, as the reference (structure.cpp:13-24); also
        arrays (n, 5) of x, y, z, elev, laf for both arguments -> array. Evaluated on the device."""
        return self._corr(p1, p2, False)

    def corr_background(self, p1, p2):
        return self._corr(p1, p2, True)

    def _descriptor_at(self, p1):
        return self._desc

    def _corr(self, p1, p2, background):
        scalar = isinstance(p1, Point) and isinstance(p2, Point)
        desc = self._descriptor_at(p1)
        a, b = self._points5(p1), self._points5(p2)
        if isinstance(p1, Point) and a.shape[0] == 1 and b.shape[0] != 1:
            a = _np.ascontiguousarray(_np.repeat(a, b.shape[0], axis=0))
        if a.shape != b.shape:
            raise ValueError("p1 and p2 must have the same shape")
        out = _np.empty(a.shape[0], _np.float32)
        _check(_libc.gpp_structure_corr_host(_C.byref(desc), _fptr(a), _fptr(b), a.shape[0], int(background), _fptr(out)))
        return float(out[0]) if scalar else out

    def localization_distance(self, point=None):
        return float(self._desc.term[0].loc_dist)

    def clone(self):
        d = _lib.StructureDesc()
        _C.memmove(_C.byref(d), _C.byref(self._desc), _C.sizeof(d))
        c = StructureFunction.__new__(type(self))
        c._desc = d
        return c


def _single(stype, h, v, w, hmax):
    d = _lib.StructureDesc()
    _check(_libc.gpp_structure_init(_C.byref(d), stype, float(h), float(v), float(w), float(hmax)))
    return d


_DEFAULT_MIN_RHO = 0.0013   # StructureFunction::default_min_rho, structure.cpp:5


class _SpatialField:
    """Scales h, v, w on the nodes of a Grid (gpp_structure_field): <Family>Structure(Grid, vec2 h, vec2 v, vec2 w, min_rho),
    structure.cpp:168-184 (Barnes), :342 (Soar), :492 (Toar), :643 (Powerlaw), :790 (Linear)."""

    def __init__(self, grid, h, v, w):
        if not isinstance(grid, Grid):
            raise ValueError("the first argument of a spatially varying structure function must be a Grid")
        shape = tuple(grid.size())
        arrays = []
        for name, a in (("h", h), ("v", v), ("w", w)):
            a = _farray(a, 2, name)
            if a.shape != shape:
                raise ValueError("Grid size not the same as scale size")
            arrays.append(a)
        self.grid = grid          # keeps the grid (and its device index) alive
        self._handle = _C.c_void_p()
        _check(_libc.gpp_structure_field_create(grid._set._handle, _fptr(arrays[0]), _fptr(arrays[1]), _fptr(arrays[2]),
                                                _C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h and _libc is not None:
            _libc.gpp_structure_field_destroy(h)
            self._handle = None


class _Family(StructureFunction):
    _TYPE = None

    def __init__(self, *args):
        self._field = None
        if len(args) > 0 and isinstance(args[0], Grid):
            if len(args) < 4 or len(args) > 5:
                raise ValueError("expected (grid, h, v, w[, min_rho])")
            grid, h, v, w = args[:4]
            min_rho = float(args[4]) if len(args) == 5 else _DEFAULT_MIN_RHO
            h2, v2, w2 = (_np.asarray(a, _np.float32) for a in (h, v, w))
            if h2.shape == (1, 1) and v2.shape == (1, 1) and w2.shape == (1, 1):
                # not spatial (structure.cpp:174-176): constant scales with the given min_rho
                d = _lib.StructureDesc()
                _check(_libc.gpp_structure_init_min_rho(_C.byref(d), self._TYPE, float(h2[0, 0]), float(v2[0, 0]), float(w2[0, 0]), min_rho))
                StructureFunction.__init__(self, d)
                return
            self._field = _SpatialField(grid, h2, v2, w2)
            self._min_rho = min_rho
            d = _lib.StructureDesc()
            d.n_terms = 1
            d.term[0].type = self._TYPE
            d.term[0].min_rho = min_rho
            # placeholder: the scales live in the field. NaN scales make every entry point that takes a plain descriptor
            # (EnSI, the device-resident forms) refuse it instead of analysing with a localization distance of 0.
            d.term[0].h = d.term[0].v = d.term[0].w = d.term[0].loc_dist = float("nan")
            StructureFunction.__init__(self, d)
            return
        h = args[0]
        v = args[1] if len(args) > 1 else 0
        w = args[2] if len(args) > 2 else 0
        hmax = args[3] if len(args) > 3 else MV
        StructureFunction.__init__(self, _single(self._TYPE, h, v, w, hmax))

    def localization_distance(self, point=None):
        if self._field is None:
            return StructureFunction.localization_distance(self, point)
        if point is None:
            raise ValueError("a spatially varying structure function needs the point (lat, lon)")
        lat, lon = (point.lat, point.lon) if hasattr(point, "lat") else (point[0], point[1])
        out = _C.c_float()
        _check(_libc.gpp_structure_field_localization_distance(self._field._handle, self._TYPE, self._min_rho, float(lat), float(lon),
                                                               _C.byref(out)))
        return float(out.value)

    def _descriptor_at(self, p1):
        """The constant-scale descriptor corr(p1, .) uses: for a spatially varying structure function the scales of the
        node nearest to p1 (structure.cpp:189-199), which needs p1's lat / lon, i.e. a Point."""
        if self._field is None:
            return self._desc
        if not isinstance(p1, Point):
            raise ValueError("corr() of a spatially varying structure function needs p1 as a Point (its lat / lon select the scales)")
        la, lo = _np.array([p1.lat], _np.float32), _np.array([p1.lon], _np.float32)
        h, v, w = (_np.empty(1, _np.float32) for _ in range(3))
        _check(_libc.gpp_structure_field_lookup_host(self._field._handle, _fptr(la), _fptr(lo), 1, _fptr(h), _fptr(v), _fptr(w)))
        d = _lib.StructureDesc()
        _check(_libc.gpp_structure_init_min_rho(_C.byref(d), self._TYPE, float(h[0]), float(v[0]), float(w[0]), self._min_rho))
        return d

    def clone(self):
        c = StructureFunction.clone(self)
        c._field = self._field
        if self._field is not None:
            c._min_rho = self._min_rho
        return c


class BarnesStructure(_Family):
    """BarnesStructure(h, v=0, w=0, hmax=MV) or BarnesStructure(grid, h, v, w, min_rho=0.0013)."""
    _TYPE = _BARNES


class CressmanStructure(StructureFunction):
    def __init__(self, h, v=0, w=0):
        StructureFunction.__init__(self, _single(_CRESSMAN, h, v, w, MV))


class SoarStructure(_Family):
    _TYPE = _SOAR


class ToarStructure(_Family):
    _TYPE = _TOAR


class PowerlawStructure(_Family):
    _TYPE = _POWERLAW


class LinearStructure(_Family):
    _TYPE = _LINEAR


def _reject_spatial(*structures):
    for st in structures:
        if getattr(st, "_field", None) is not None:
            raise NotImplementedOnDevice("spatially varying structure functions cannot be nested in MultipleStructure / CrossValidation on the device")


class MultipleStructure(StructureFunction):
    def __init__(self, structure_h, structure_v, structure_w):
        _reject_spatial(structure_h, structure_v, structure_w)
        d = _lib.StructureDesc()
        _check(_libc.gpp_structure_multiple(_C.byref(d), _C.byref(structure_h._desc), _C.byref(structure_v._desc),
                                            _C.byref(structure_w._desc)))
        StructureFunction.__init__(self, d)


class CrossValidation(StructureFunction):
    def __init__(self, structure, dist=MV):
        _reject_spatial(structure)
        d = _lib.StructureDesc()
        _check(_libc.gpp_structure_cross_validation(_C.byref(d), _C.byref(structure._desc), float(dist)))
        StructureFunction.__init__(self, d)


# ---------------------------------------------------------------------------------------------------------
def _is_grid(obj):
    return isinstance(obj, Grid)


def _check_size(cond, msg):
    if not cond:
        raise ValueError(msg)


def optimal_interpolation(bgrid, background, points, pobs, pratios, pbackground, structure, max_points,
                          allow_extrapolation=True):
    """gridpp::optimal_interpolation, oi.cpp:26-136 (Grid and Points overloads)."""
    out = _oi(bgrid, background, None, points, pobs, pratios, pbackground, None, structure, max_points, allow_extrapolation, False)
    return out


def optimal_interpolation_multi_gpu(bgrid, background, points, pobs, pratios, pbackground, structure, max_points,
                                    allow_extrapolation=True, n_devices=0):
    """gridpp::optimal_interpolation over every visible GPU from ONE process (gpp_optimal_interpolation_multi_gpu_host): the rows of
    the grid are split into one block per device. Bit-identical to the single-device call. n_devices = 0: all devices."""
    return _oi(bgrid, background, None, points, pobs, pratios, pbackground, None, structure, max_points, allow_extrapolation, False,
               n_devices=n_devices)


def optimal_interpolation_full(bgrid, background, bvariance, points, obs, obs_variance, background_at_points,
                               bvariance_at_points, structure, max_points, allow_extrapolation=True):
    """gridpp::optimal_interpolation_full, oi.cpp:138-412. Returns (analysis, analysis_variance)."""
    return _oi(bgrid, background, bvariance, points, obs, obs_variance, background_at_points, bvariance_at_points, structure,
               max_points, allow_extrapolation, True)


def _oi(bgrid, background, bvariance, points, pobs, obs_variance, pbackground, bvariance_at_points, structure, max_points,
        allow_extrapolation, full, n_devices=None):
    if max_points < 0:
        raise ValueError("max_points must be >= 0")
    if not isinstance(points, Points):
        raise ValueError("points must be a Points object")
    if bgrid.get_coordinate_type() != points.get_coordinate_type():
        raise ValueError("Both background and observations points must be of same coordinate type (lat/lon or x/y)")
    grid = _is_grid(bgrid)
    ndim = 2 if grid else 1
    bg = _farray(background, ndim, "background")
    shape = tuple(bgrid.size()) if grid else (bgrid.size(),)
    _check_size(bg.shape == shape, "input field %s is not the same size as the grid %s" % (bg.shape, shape))
    bvar = None
    if bvariance is not None:
        bvar = _farray(bvariance, ndim, "bvariance")
        _check_size(bvar.shape == shape, "Input bvariance %s is not the same size as the grid %s" % (bvar.shape, shape))
    nS = points.size()
    obs = _farray(pobs, 1, "pobs")
    ovar = _farray(obs_variance, 1, "obs_variance")
    pbg = _farray(pbackground, 1, "pbackground")
    _check_size(obs.size == nS, "Observations (%d) and points (%d) size mismatch" % (obs.size, nS))
    _check_size(ovar.size == nS, "Ratios (%d) and points (%d) size mismatch" % (ovar.size, nS))
    _check_size(pbg.size == nS, "Background (%d) and points (%d) size mismatch" % (pbg.size, nS))
    pbvar = None
    if bvariance_at_points is not None:
        pbvar = _farray(bvariance_at_points, 1, "bvariance_at_points")
        _check_size(pbvar.size == nS, "Background variance (%d) and points (%d) size mismatch" % (pbvar.size, nS))
    out = _np.empty(shape, _np.float32)
    var = _np.empty(shape, _np.float32) if full else None
    field = getattr(structure, "_field", None)
    if field is not None:
        _check(_libc.gpp_optimal_interpolation_spatial_host(bgrid._set._handle, _fptr(bg), _fptr(bvar), points._set._handle, _fptr(obs),
                                                            _fptr(ovar), _fptr(pbg), _fptr(pbvar), int(structure._TYPE), field._handle,
                                                            float(structure._min_rho), int(max_points), int(bool(allow_extrapolation)),
                                                            _fptr(out), _fptr(var)))
        return (out, var) if full else out
    if n_devices is not None:
        _check(_libc.gpp_optimal_interpolation_multi_gpu_host(int(n_devices), bgrid._set._handle, _fptr(bg), _fptr(bvar), points._set._handle,
                                                              _fptr(obs), _fptr(ovar), _fptr(pbg), _fptr(pbvar), _C.byref(structure._desc),
                                                              int(max_points), int(bool(allow_extrapolation)), _fptr(out), _fptr(var)))
        return (out, var) if full else out
    _check(_libc.gpp_optimal_interpolation_host(bgrid._set._handle, _fptr(bg), _fptr(bvar), points._set._handle, _fptr(obs),
                                                _fptr(ovar), _fptr(pbg), _fptr(pbvar), _C.byref(structure._desc),
                                                int(max_points), int(bool(allow_extrapolation)), _fptr(out), _fptr(var)))
    return (out, var) if full else out


def optimal_interpolation_ensi(bgrid, background, points, pobs, psigmas, pbackground, structure, max_points,
                               allow_extrapolation=True):
    """gridpp::optimal_interpolation_ensi, oi_ensi.cpp:33-568. background is (Y, X, E) for a Grid or (N, E) for
    Points; pbackground is (S, E)."""
    if max_points < 0:
        raise ValueError("max_points must be >= 0")
    grid = _is_grid(bgrid)
    bg = _farray(background, 3 if grid else 2, "background")
    nS = points.size()
    if nS == 0:
        return bg.copy()   # oi_ensi.cpp:49-51,137-139
    if bgrid.get_coordinate_type() != points.get_coordinate_type():
        raise ValueError("Both background and observations points must be of same coorindate type (lat/lon or x/y)")
    shape = tuple(bgrid.size()) if grid else (bgrid.size(),)
    if grid and (shape[0] == 0 or shape[1] == 0):
        raise ValueError("Grid size (%d,%d) cannot be zero" % shape)
    _check_size(bg.shape[:-1] == shape, "Input field is not the same size as the grid")
    nE = bg.shape[-1]
    obs = _farray(pobs, 1, "pobs")
    sig = _farray(psigmas, 1, "psigmas")
    pbg = _farray(pbackground, 2, "pbackground")
    _check_size(obs.size == nS, "Observations and points exception mismatch")
    _check_size(sig.size == nS, "Sigmas and points size mismatch")
    _check_size(pbg.shape[0] == nS, "Background and points size mismatch")
    _check_size(pbg.shape[1] == nE, "Background and points ensemble size mismatch")
    out = _np.empty(bg.shape, _np.float32)
    skipped = _C.c_int()
    _check(_libc.gpp_optimal_interpolation_ensi_host(bgrid._set._handle, _fptr(bg), nE, points._set._handle, _fptr(obs), _fptr(sig),
                                                     _fptr(pbg), _C.byref(structure._desc), int(max_points),
                                                     int(bool(allow_extrapolation)), _fptr(out), _C.byref(skipped)))
    if skipped.value > 0:   # oi_ensi.cpp:557-561
        print("Warning: Condition number error in %d points. Using raw values in those points." % skipped.value)
    return out


def _ensi_multi(kind, bgrid, bratios, background, background_corr, points, pobs, pratios, pbackground, pbackground_corr, structure,
                max_points, allow_extrapolation):
    """Shared host side of the three "multi" variants (oi_ensi_multi.cpp:33-327 Grid overloads, :329-1311 Points overloads)."""
    if max_points < 0:
        raise ValueError("max_points must be >= 0")
    grid = _is_grid(bgrid)
    bg = _farray(background, 3 if grid else 2, "background")
    nS = points.size()
    if nS == 0:
        return bg.copy()
    shape = tuple(bgrid.size()) if grid else (bgrid.size(),)
    if grid and (shape[0] == 0 or shape[1] == 0):
        raise ValueError("Grid size (%d,%d) cannot be zero" % shape)
    if bgrid.get_coordinate_type() != points.get_coordinate_type():
        raise ValueError("Both background and observations points must be of same coorindate type (lat/lon or x/y)")
    _check_size(bg.shape[:-1] == shape, "Input background field is not the same size as the grid")
    nE = bg.shape[-1]
    br = _farray(bratios, 2 if grid else 1, "bratios")
    _check_size(br.shape == shape, "Bratios and grid size mismatch")
    obs = _farray(pobs, 1 if kind == "utem" else 2, "pobs")
    pr = _farray(pratios, 1, "pratios")
    pbg = _farray(pbackground, 2, "pbackground")
    _check_size(obs.shape[0] == nS and (kind == "utem" or obs.shape[1] == nE), "Observations and points exception mismatch")
    _check_size(pr.size == nS, "Pratios and points size mismatch")
    _check_size(pbg.shape == (nS, nE), "Background and points size mismatch")
    bgc = pbgc = None
    if kind != "ebesc":
        bgc = _farray(background_corr, 3 if grid else 2, "background_corr")
        pbgc = _farray(pbackground_corr, 2, "pbackground_corr")
        _check_size(bgc.shape == bg.shape, "Input background_corr field is not the same size as the grid")
        _check_size(pbgc.shape == (nS, nE), "Background_corr and points size mismatch")
    out = _np.empty(bg.shape, _np.float32)
    common = (_C.byref(structure._desc), int(max_points), int(bool(allow_extrapolation)), _fptr(out))
    if kind == "ebesc":
        _check(_libc.gpp_optimal_interpolation_ensi_multi_ebesc_host(bgrid._set._handle, _fptr(br), _fptr(bg), nE, points._set._handle, _fptr(obs),
                                                                     _fptr(pr), _fptr(pbg), *common))
    elif kind == "ebe":
        _check(_libc.gpp_optimal_interpolation_ensi_multi_ebe_host(bgrid._set._handle, _fptr(br), _fptr(bg), _fptr(bgc), nE, points._set._handle,
                                                                   _fptr(obs), _fptr(pr), _fptr(pbg), _fptr(pbgc), *common))
    else:
        skipped = _C.c_int()
        _check(_libc.gpp_optimal_interpolation_ensi_multi_utem_host(bgrid._set._handle, _fptr(br), _fptr(bg), _fptr(bgc), nE, points._set._handle,
                                                                    _fptr(obs), _fptr(pr), _fptr(pbg), _fptr(pbgc), *common, _C.byref(skipped)))
        if skipped.value > 0:   # oi_ensi_multi.cpp:1300-1304
            print("Warning: Condition number error in %d points. Using raw values in those points." % skipped.value)
    return out


def optimal_interpolation_ensi_multi_ebe(bgrid, bratios, background, background_corr, points, pobs, pratios, pbackground,
                                         pbackground_corr, structure, max_points, allow_extrapolation=True):
    """gridpp::optimal_interpolation_ensi_multi_ebe, oi_ensi_multi.cpp:34-135 (Grid) and :329-627 (Points): member-by-member
    increments with ensemble-based correlations. background* is (Y, X, E) / (L, E); pobs, pbackground* are (S, E)."""
    return _ensi_multi("ebe", bgrid, bratios, background, background_corr, points, pobs, pratios, pbackground, pbackground_corr, structure,
                       max_points, allow_extrapolation)


def optimal_interpolation_ensi_multi_ebesc(bgrid, bratios, background, points, pobs, pratios, pbackground, structure, max_points,
                                           allow_extrapolation=True):
    """gridpp::optimal_interpolation_ensi_multi_ebesc, oi_ensi_multi.cpp:137-224 (Grid) and :630-859 (Points): member-by-member
    increments with static correlations."""
    return _ensi_multi("ebesc", bgrid, bratios, background, None, points, pobs, pratios, pbackground, None, structure, max_points,
                       allow_extrapolation)


def optimal_interpolation_ensi_multi_utem(bgrid, bratios, background, background_corr, points, pobs, pratios, pbackground,
                                          pbackground_corr, structure, max_points, allow_extrapolation=True):
    """gridpp::optimal_interpolation_ensi_multi_utem, oi_ensi_multi.cpp:226-327 (Grid) and :862-1311 (Points): the ensemble transform of EnSI computed from
    the *_corr ensembles, applied to the mean and spread of background. pobs is (S,)."""
    return _ensi_multi("utem", bgrid, bratios, background, background_corr, points, pobs, pratios, pbackground, pbackground_corr, structure,
                       max_points, allow_extrapolation)


def staticcorr_points(points, knots, structure, max_points):
    """gridpp::staticcorr_points, corr_points.cpp:26-131 -> (L, K): row l holds structure.corr_background(point l, knot) for the
    (at most max_points best) knots within the localization radius, 0 elsewhere."""
    if max_points < 0:
        raise ValueError("max_points must be >= 0")
    if points.get_coordinate_type() != knots.get_coordinate_type():
        raise ValueError("Both background grid and observations points must be of same coordinate type (lat/lon or x/y)")
    out = _np.zeros((points.size(), knots.size()), _np.float32)
    if out.size:
        _check(_libc.gpp_staticcorr_points_host(points._set._handle, knots._set._handle, _C.byref(structure._desc), int(max_points), _fptr(out)))
    return out


# ---------------------------------------------------------------------------------------------------------
def neighbourhood(input, halfwidth, statistic):
    """gridpp::neighbourhood(vec2, halfwidth, statistic), neighbourhood.cpp:28-242."""
    field = _np.asarray(input, dtype=_np.float32)
    if field.ndim == 3:
        # gridpp::neighbourhood(vec3, ...), neighbourhood.cpp:12-27: members reduced with calc_statistic first
        field = _farray(field, 3, "input")
        if halfwidth < 0:
            raise ValueError("Half width must be > 0")
        if statistic == Quantile:
            raise ValueError("Use neighbourhood_quantile for computing neighbourhood quantiles")
        ny, nx, ne = field.shape
        if ny == 0 or nx == 0 or ne == 0:
            return _np.zeros((0, 0), _np.float32)
        out = _np.empty((ny, nx), _np.float32)
        _check(_libc.gpp_neighbourhood_ens_host(_fptr(field), ny, nx, ne, int(halfwidth), int(statistic), _fptr(out)))
        return out
    field = _farray(field, 2, "input")
    if halfwidth < 0:
        raise ValueError("Half width must be > 0")
    if statistic == Quantile:
        raise ValueError("Use neighbourhood_quantile for computing neighbourhood quantiles")
    if field.shape[0] == 0 or field.shape[1] == 0:
        return _np.zeros((0, 0), _np.float32)
    out = _np.empty(field.shape, _np.float32)
    _check(_libc.gpp_neighbourhood_host(_fptr(field), field.shape[0], field.shape[1], int(halfwidth), int(statistic), _fptr(out)))
    return out


def neighbourhood_search(array, search_array, halfwidth, search_target_min, search_target_max, search_delta, apply_array=None):
    """gridpp::neighbourhood_search, neighbourhood_search.cpp:7-113: the mean of `array` over the window cells whose
    `search_array` value is inside [search_target_min, search_target_max]; failing that the value at the cell closest to that
    range (at least search_delta away from the centre's own search value); failing that the input value."""
    a = _farray(array, 2, "array")
    sa = _farray(search_array, 2, "search_array")
    if search_target_min > search_target_max:
        raise ValueError("Search_target_min must be smaller than search_target_max")
    if halfwidth < 0:
        raise ValueError("halfwidth must be positive")
    if sa.shape != a.shape:
        raise ValueError("search_array must either be the same size as array")
    ap = None
    if apply_array is not None:
        ap = _np.ascontiguousarray(apply_array, dtype=_np.int32)
        if ap.ndim != 2:
            raise ValueError("apply_array must have 2 dimensions")
        if ap.shape[0] > 1 and ap.shape != a.shape:
            raise ValueError("apply_array must either be empty or same size as array")
        if ap.size == 0:
            ap = None
        elif ap.shape != a.shape:
            raise ValueError("apply_array must either be empty or same size as array")
    out = _np.empty(a.shape, _np.float32)
    if a.size:
        _check(_libc.gpp_neighbourhood_search_host(_fptr(a), _fptr(sa), a.shape[0], a.shape[1], int(halfwidth), float(search_target_min),
                                                   float(search_target_max), float(search_delta),
                                                   ap.ctypes.data_as(_C.POINTER(_C.c_int)) if ap is not None else None, _fptr(out)))
    return out


def calc_gradient(base, values, gradient_type, halfwidth, min_num=2, min_range=MV, default_gradient=0):
    """gridpp::calc_gradient, calc_gradient.cpp:6-126: the gradient of `values` with respect to `base` in every neighbourhood, from
    the cells with the lowest and highest base (MinMax) or by linear regression over the window (LinearRegression)."""
    b = _farray(base, 2, "base")
    v = _farray(values, 2, "values")
    if halfwidth <= 0:
        raise ValueError("Halwidth cannot be <= 0; must be positive integer")
    if is_valid(min_range) and min_range < 0:
        raise ValueError("min_range must be >= 0")
    if min_num < 0:
        raise ValueError("num_min must be >= 0")
    if b.shape[0] == 0:
        raise ValueError("base input has no size")
    if b.shape != v.shape:
        raise ValueError("base is not the same size as values")
    out = _np.empty(b.shape, _np.float32)
    if b.size:
        _check(_libc.gpp_calc_gradient_host(_fptr(b), _fptr(v), b.shape[0], b.shape[1], int(gradient_type), int(halfwidth), int(min_num),
                                            float(min_range), float(default_gradient), _fptr(out)))
    return out


def _brute_force(input, halfwidth, statistic, quantile):
    field = _np.asarray(input, dtype=_np.float32)
    if field.ndim not in (2, 3):
        raise ValueError("input must have 2 or 3 dimensions")
    field = _farray(field, field.ndim, "input")
    if halfwidth < 0:
        raise ValueError("Half width must be > 0")
    if 0 in field.shape:
        return _np.zeros((0, 0), _np.float32)
    ny, nx = field.shape[:2]
    ne = field.shape[2] if field.ndim == 3 else 1
    out = _np.empty((ny, nx), _np.float32)
    _check(_libc.gpp_neighbourhood_brute_force_host(_fptr(field), ny, nx, ne, int(halfwidth), int(statistic), float(quantile), _fptr(out)))
    return out


def neighbourhood_brute_force(input, halfwidth, statistic):
    """gridpp::neighbourhood_brute_force(vec2 | vec3, halfwidth, statistic), neighbourhood.cpp:528-533."""
    return _brute_force(input, halfwidth, statistic, 0.0)


def neighbourhood_quantile(input, quantile, halfwidth):
    """gridpp::neighbourhood_quantile(vec2 | vec3, quantile, halfwidth), neighbourhood.cpp:534-539: the exact quantile."""
    return _brute_force(input, halfwidth, Quantile, quantile)


def neighbourhood_ens(input, halfwidth, statistic):
    future_deprecation_warning("neighbourhood_ens", "neighbourhood")                 # neighbourhood.cpp:541-544
    return neighbourhood(input, halfwidth, statistic)


def neighbourhood_quantile_ens(input, quantile, halfwidth):
    future_deprecation_warning("neighbourhood_quantile_ens", "neighbourhood_quantile")
    return neighbourhood_quantile(input, quantile, halfwidth)


def neighbourhood_quantile_ens_fast(input, quantile, halfwidth, thresholds):
    future_deprecation_warning("neighbourhood_quantile_ens_fast", "neighbourhood_quantile_fast")
    return neighbourhood_quantile_fast(input, quantile, halfwidth, thresholds)


def neighbourhood_quantile_fast(input, quantile, halfwidth, thresholds):
    """gridpp::neighbourhood_quantile_fast(vec2, float | vec2, halfwidth, thresholds), neighbourhood.cpp:296-409."""
    field = _np.asarray(input, dtype=_np.float32)
    ens = field.ndim == 3     # gridpp::neighbourhood_quantile_fast(vec3, ...), neighbourhood.cpp:411-527
    field = _farray(field, 3 if ens else 2, "input")
    thr = _farray(thresholds, 1, "thresholds")
    if halfwidth < 0:
        raise ValueError("Half width must be > 0")
    if 0 in field.shape:
        return _np.zeros((0, 0), _np.float32)
    qf = None
    q = float("nan")
    if _np.ndim(quantile) == 0:
        q = float(quantile)
    else:
        qf = _farray(quantile, 2, "quantile")
        if qf.shape == (1, 1):
            q, qf = float(qf[0, 0]), None
        elif qf.shape != field.shape[:2]:
            raise ValueError("Quantile must have the same Y, X size as input, or have size (1, 1)")
    out = _np.empty(field.shape[:2], _np.float32)
    if ens:
        _check(_libc.gpp_neighbourhood_quantile_fast_ens_host(_fptr(field), field.shape[0], field.shape[1], field.shape[2], q, _fptr(qf),
                                                              int(halfwidth), _fptr(thr), thr.size, _fptr(out)))
        return out
    _check(_libc.gpp_neighbourhood_quantile_fast_host(_fptr(field), field.shape[0], field.shape[1], q, _fptr(qf), int(halfwidth),
                                                      _fptr(thr), thr.size, _fptr(out)))
    return out


def get_neighbourhood_thresholds(input, num_thresholds):
    """gridpp::get_neighbourhood_thresholds(vec2 | vec3, num_thresholds), neighbourhood.cpp:243-295."""
    field = _np.asarray(input, dtype=_np.float32)
    if field.size == 0 and num_thresholds > 0:
        return _np.zeros(0, _np.float32)          # neighbourhood.cpp:247-248 (an empty list arrives as an empty vec2)
    if field.ndim not in (2, 3):
        raise ValueError("input must have 2 or 3 dimensions")
    if num_thresholds <= 0:
        raise ValueError("num_thresholds must be > 0")
    if field.size == 0:
        return _np.zeros(0, _np.float32)
    flat = _np.ascontiguousarray(field.ravel())
    out = _np.empty(int(num_thresholds), _np.float32)
    n = _C.c_int()
    _check(_libc.gpp_get_neighbourhood_thresholds_host(_fptr(flat), flat.size, int(num_thresholds), _fptr(out), _C.byref(n)))
    return out[:n.value].copy()


# ---------------------------------------------------------------------------------------------------------
# consumers of the point index: gridding / gridding_nearest / count / distance / fill / fill_missing / doping
def _out_shape(obj):
    # gridding.cpp:14-15, count.cpp:20-21, distance.cpp:33: the output is sized by Grid::size(), (0, 0) for a grid without nodes
    return tuple(obj.size()) if _is_grid(obj) else (obj.size(),)


def gridding(grid, points, values, radius, min_num, statistic):
    """gridpp::gridding(Grid | Points, Points, values, radius, min_num, statistic), gridding.cpp:6-61."""
    vals = _farray(values, 1, "values")
    _check_size(points.size() == vals.size, "Points size is not the same as values")
    if not is_valid(radius) or radius < 0:
        raise ValueError("radius must be >= 0")
    if min_num < 0:
        raise ValueError("min_num must be >= 0")
    out = _np.empty(grid._set.n, _np.float32)
    _check(_libc.gpp_gridding_host(grid._set._handle, points._set._handle, _fptr(vals), float(radius), int(min_num), int(statistic), _fptr(out)))
    return out.reshape(_out_shape(grid))


def gridding_nearest(grid, points, values, min_num, statistic):
    """gridpp::gridding_nearest(Grid | Points, Points, values, min_num, statistic), gridding.cpp:63-131."""
    vals = _farray(values, 1, "values")
    _check_size(points.size() == vals.size, "Points size is not the same as values")
    if min_num < 0:
        raise ValueError("min_num must be >= 0")
    out = _np.empty(grid._set.n, _np.float32)
    _check(_libc.gpp_gridding_nearest_host(grid._set._handle, points._set._handle, _fptr(vals), int(min_num), int(statistic), _fptr(out)))
    return out.reshape(_out_shape(grid))


def count(iobj, oobj, radius):
    """gridpp::count (count.cpp:6-66): the number of points / nodes of `iobj` within `radius` of every location of `oobj`."""
    out = _np.empty(oobj._set.n, _np.float32)
    _check(_libc.gpp_count_host(iobj._set._handle, oobj._set._handle, float(radius), _fptr(out)))
    return out.reshape(_out_shape(oobj))


def distance(iobj, oobj, num=1):
    """gridpp::distance (distance.cpp:6-120): the distance from every location of `oobj` to its num-th closest point of `iobj`."""
    if iobj.get_coordinate_type() != oobj.get_coordinate_type():
        raise ValueError("Incompatible coordinate types")
    # distance.cpp:21 (Grid, Points) and :111 (Points, Points) pass the output location first; :52, :83 the input point
    query_first = not _is_grid(oobj)
    out = _np.empty(oobj._set.n, _np.float32)
    _check(_libc.gpp_distance_host(iobj._set._handle, oobj._set._handle, int(num), int(query_first), _fptr(out)))
    return out.reshape(_out_shape(oobj))


def fill(igrid, input, points, radii, value, outside):
    """gridpp::fill, fill.cpp:6-43."""
    field = _farray(input, 2, "input")
    _check_size(field.shape == tuple(igrid._stored_shape), "Grid size is not the same as values")
    rad = _farray(radii, 1, "radii")
    _check_size(points.size() == rad.size, "Points size is not the same as radii size")
    if (rad < 0).any():
        raise ValueError("All radius sizes must be 0 or greater")
    out = _np.empty(field.shape, _np.float32)
    _check(_libc.gpp_fill_host(igrid._set._handle, _fptr(field), points._set._handle, _fptr(rad), float(value), int(bool(outside)), _fptr(out)))
    return out


def fill_missing(values):
    """gridpp::fill_missing, fill.cpp:44-134."""
    field = _farray(values, 2, "values")
    out = _np.empty(field.shape, _np.float32)
    if field.size:
        _check(_libc.gpp_fill_missing_host(_fptr(field), field.shape[0], field.shape[1], _fptr(out)))
    return out


def _doping_checks(igrid, background, points, observations, extent, what):
    field = _farray(background, 2, "background")
    _check_size(field.shape == tuple(igrid._stored_shape), "Grid size is not the same as observations")
    obs = _farray(observations, 1, "observations")
    _check_size(points.size() == obs.size, "Points size is not the same as observations size")
    _check_size(points.size() == len(extent), "Points size is not the same as %s size" % what)
    return field, obs


def doping_square(igrid, background, points, observations, halfwidths, max_elev_diff=MV):
    """gridpp::doping_square, doping.cpp:5-51."""
    hw = _np.ascontiguousarray(_np.asarray(halfwidths, dtype=_np.int32).ravel())
    field, obs = _doping_checks(igrid, background, points, observations, hw, "halfwidth")
    if is_valid(max_elev_diff) and max_elev_diff < 0:
        raise ValueError("max_elev_diff must be greater than or equal to 0")
    if (hw < 0).any():
        raise ValueError("All halfwidth must be greater than or equal to 0")
    out = _np.empty(field.shape, _np.float32)
    if field.size:
        _check(_libc.gpp_doping_square_host(igrid._set._handle, _fptr(field), points._set._handle, _fptr(obs), hw.ctypes.data_as(_lib.ip),
                                            float(max_elev_diff), _fptr(out)))
    return out


def doping_circle(igrid, background, points, observations, radii, max_elev_diff=MV):
    """gridpp::doping_circle, doping.cpp:52-93."""
    rad = _farray(radii, 1, "radii")
    field, obs = _doping_checks(igrid, background, points, observations, rad, "radii")
    if is_valid(max_elev_diff) and max_elev_diff < 0:
        raise ValueError("max_elev_diff must be greater than or equal to 0")
    if (rad < 0).any():
        raise ValueError("radii must be greater than or equal to 0")
    out = _np.empty(field.shape, _np.float32)
    if field.size:
        _check(_libc.gpp_doping_circle_host(igrid._set._handle, _fptr(field), points._set._handle, _fptr(obs), _fptr(rad), float(max_elev_diff),
                                            _fptr(out)))
    return out


# ---------------------------------------------------------------------------------------------------------
_STATISTIC_NAMES = {"mean": Mean, "min": Min, "max": Max, "median": Median, "quantile": Quantile, "std": Std, "sum": Sum,
                    "count": Count, "randomchoice": RandomChoice}


def get_statistic(name):
    """gridpp::get_statistic, gridpp.cpp:11-44 (note: no "variance" entry in the reference)."""
    return _STATISTIC_NAMES.get(name, Unknown)


def calc_statistic(array, statistic):
    """gridpp::calc_statistic: (vec, statistic) -> float (util.cpp:19-110), (vec2, statistic) -> vec (:208-215)."""
    a = _np.asarray(array, dtype=_np.float32)
    if a.ndim not in (1, 2):
        raise ValueError("array must be 1- or 2-dimensional")
    rows = _np.ascontiguousarray(a.reshape(1, -1) if a.ndim == 1 else a)
    out = _np.empty(rows.shape[0], _np.float32)
    if rows.shape[0] > 0:
        _check(_libc.gpp_calc_statistic_host(_fptr(rows), rows.shape[0], rows.shape[1], int(statistic), _fptr(out)))
    return float(out[0]) if a.ndim == 1 else out


def calc_quantile(array, quantile=MV):
    """gridpp::calc_quantile: (vec, q) -> float (util.cpp:111-178), (vec2, q) -> vec (:179-186), (vec3, vec2 q) -> vec2 (:187-207)."""
    a = _np.asarray(array, dtype=_np.float32)
    if a.ndim == 3:
        q = _farray(quantile, 2, "quantile")
        if a.shape[:2] != q.shape:
            raise ValueError("Dimension mismatch between array and quantile")
        if a.shape[0] == 0 or a.shape[1] == 0:
            return _np.zeros((0, 0), _np.float32)               # util.cpp:190-195: an empty vec2
        if a.shape[2] == 0:
            return _np.full(a.shape[:2], MV, _np.float32)
        rows = _np.ascontiguousarray(a.reshape(-1, a.shape[2]))
        out = _np.empty(rows.shape[0], _np.float32)
        _check(_libc.gpp_calc_quantile_host(_fptr(rows), rows.shape[0], rows.shape[1], float("nan"), _fptr(_np.ascontiguousarray(q.ravel())), _fptr(out)))
        return out.reshape(a.shape[:2])
    if a.ndim not in (1, 2):
        raise ValueError("array must be 1-, 2- or 3-dimensional")
    q = float(quantile)
    if q < 0 or q > 1:
        raise ValueError("calc_quantile: Quantile must be between 0 and 1 inclusive")
    rows = _np.ascontiguousarray(a.reshape(1, -1) if a.ndim == 1 else a)
    out = _np.empty(rows.shape[0], _np.float32)
    if rows.shape[0] > 0:
        _check(_libc.gpp_calc_quantile_host(_fptr(rows), rows.shape[0], rows.shape[1], q, None, _fptr(out)))
    return float(out[0]) if a.ndim == 1 else out


def interpolate(x, iX, iY):
    """gridpp::interpolate: (float, vec, vec) -> float (util.cpp:377-414), (vec, vec, vec) -> vec (:415-431)."""
    ix, iy = _farray(iX, 1, "iX"), _farray(iY, 1, "iY")
    if ix.size != iy.size:
        raise ValueError("Dimension mismatch. Cannot interpolate.")
    xs = _np.asarray(x, dtype=_np.float32)
    scalar = xs.ndim == 0
    flat = _np.ascontiguousarray(xs.reshape(-1))
    out = _np.empty(flat.size, _np.float32)
    if flat.size > 0:
        _check(_libc.gpp_interpolate_host(_fptr(flat), flat.size, _fptr(ix), _fptr(iy), ix.size, _fptr(out)))
    return float(out[0]) if scalar else out


def calc_even_quantiles(values, num):
    """gridpp::calc_even_quantiles, util.cpp:261-338: evenly spaced quantiles among the distinct values (the selection
    gpp_get_neighbourhood_thresholds_host makes after its device sort)."""
    v = _farray(values, 1, "values")
    if int(num) == 0 or v.size == 0:
        return _np.zeros(0, _np.float32)
    out = _np.empty(int(num), _np.float32)
    n = _C.c_int()
    _check(_libc.gpp_get_neighbourhood_thresholds_host(_fptr(v), v.size, int(num), _fptr(out), _C.byref(n)))
    return out[:n.value].copy()


def get_lower_index(x, values):
    """gridpp::get_lower_index, util.cpp:339-357: last index whose value is <= x (first of an exact match)."""
    index = None
    for i, c in enumerate(_np.asarray(values, _np.float32)):
        if not is_valid(float(c)):
            continue
        if c < x:
            index = i
        elif c == x:
            index = i
            break
        else:
            break
    if index is None:
        raise RuntimeError("get_lower_index: no valid value at or below x")   # the reference returns an undefined int
    return index


def get_upper_index(x, values):
    """gridpp::get_upper_index, util.cpp:358-376."""
    index = None
    vals = _np.asarray(values, _np.float32)
    for i in range(len(vals) - 1, -1, -1):
        c = vals[i]
        if not is_valid(float(c)):
            continue
        if c > x:
            index = i
        elif c == x:
            index = i
            break
        else:
            break
    if index is None:
        raise RuntimeError("get_upper_index: no valid value at or above x")
    return index


def num_missing_values(array):
    """gridpp::num_missing_values, util.cpp:216-224"""
    return int((~_np.isfinite(_np.asarray(array, _np.float32))).sum())


def is_valid_lat(lat, type):
    """util.cpp:617-621"""
    if type == Cartesian:
        return is_valid(lat)
    return is_valid(lat) and -90.001 <= lat <= 90.001


def is_valid_lon(lon, type):
    """util.cpp:622-624"""
    return is_valid(lon)


def compatible_size(a, b):
    """gridpp::compatible_size, util.cpp:421-474 (Grid | Points | array against an array)."""
    bb = _np.asarray(b)
    if isinstance(a, Grid):
        shape = tuple(a.size())
        if bb.ndim == 2:
            return bb.shape == shape
        return bb.size == 0 or bb.shape[1:] == shape        # (T, Y, X)
    if isinstance(a, Points):
        if bb.ndim == 1:
            return a.size() == bb.shape[0]
        return bb.shape[0] == 0 or a.size() == bb.shape[1]   # (T, N)
    aa = _np.asarray(a)
    if aa.ndim == 2 and bb.ndim == 3:
        return aa.shape == bb.shape[:2] or (aa.size == 0 and bb.size == 0)
    return aa.shape == bb.shape


_debug_level = 0


def set_debug_level(level):
    global _debug_level
    _debug_level = int(level)


def get_debug_level():
    return _debug_level


def debug(string):
    print(string)


def warning(string):
    print("Warning: " + string)


def error(string):
    print("Error: " + string)
    raise RuntimeError(string)


def future_deprecation_warning(function, other=""):
    print("Future deprecation warning: %s will be deprecated%s" % (function, (", use %s instead." % other) if other != "" else "."))


def clock():
    import time
    return time.time()


# ---------------------------------------------------------------------------------------------------------
def nearest(igrid, ogrid, ivalues):
    """gridpp::nearest, nearest.cpp:7-222 (all eight Grid/Points x Grid/Points x 2-D/3-D overloads)."""
    in_grid, out_grid = _is_grid(igrid), _is_grid(ogrid)
    ishape = tuple(igrid.size()) if in_grid else (igrid.size(),)
    oshape = tuple(ogrid._stored_shape) if out_grid else (ogrid.size(),)   # nearest.cpp:79-82: the output has the shape of get_lats()
    base = 2 if in_grid else 1
    values = _np.asarray(ivalues, dtype=_np.float32)
    multi = values.ndim == base + 1
    if values.ndim not in (base, base + 1):
        raise ValueError("ivalues has the wrong number of dimensions")
    if multi:
        _check_size(values.shape[1:] == ishape or values.size == 0, "Grid size is not the same as values")
        flat = _np.ascontiguousarray(values.reshape(values.shape[0], -1))
        nf = values.shape[0]
    else:
        _check_size(values.shape == ishape, "Grid size is not the same as values" if in_grid else "Points size is not the same as values")
        flat = _np.ascontiguousarray(values.reshape(1, -1))
        nf = 1
    nq = int(_np.prod(oshape))
    out = _np.empty((nf, nq), _np.float32)
    _check(_libc.gpp_nearest_host(igrid._set._handle, _fptr(ogrid._set._lats), _fptr(ogrid._set._lons), nq, _fptr(flat), nf,
                                  _fptr(out)))
    return out.reshape((nf,) + oshape) if multi else out.reshape(oshape)


# ---------------------------------------------------------------------------------------------------------
# Host-side helpers of the reference's public surface that involve no device work.
def init_vec2(Y, X, value=MV):
    """gridpp::init_vec2, util.cpp (gridpp.h:1563)"""
    return _np.full((int(Y), int(X)), value, _np.float32)


def init_ivec2(Y, X, value):
    return _np.full((int(Y), int(X)), value, _np.int32)


def init_vec3(Y, X, E, value=MV):
    return _np.full((int(Y), int(X), int(E)), value, _np.float32)


def init_ivec3(Y, X, E, value):
    return _np.full((int(Y), int(X), int(E)), value, _np.int32)


def point_in_rectangle(A, B, C, D, m):
    """gridpp::point_in_rectangle, util.cpp:562-581: float arithmetic on the corners' lat / lon, both orientations accepted."""
    f = _np.float32

    def side(p1, p2, q):
        lon, lat = f(f(p2.lon) - f(p1.lon)), f(f(-1) * f(f(p2.lat) - f(p1.lat)))       # vect2d, :563-568
        c = f(f(-1) * f(f(lat * f(p1.lon)) + f(lon * f(p1.lat))))
        return f(f(f(lat * f(q.lon)) + f(lon * f(q.lat))) + c)
    d1, d2, d3, d4 = side(A, B, m), side(A, D, m), side(B, C, m), side(C, D, m)
    opt1 = 0 >= d1 and 0 >= d4 and 0 <= d2 and 0 >= d3
    opt2 = 0 <= d1 and 0 <= d4 and 0 >= d2 and 0 <= d3
    return bool(opt1 or opt2)


# The marshalling self-tests of the reference's SWIG module (src/api/swig.cpp:6-104, tests/test_swig.py): they pin down what
# goes in (lists, tuples, arrays of any dtype; a wrong number of dimensions is an error; zero-length inputs of any rank are
# accepted) and what comes out (float32 / int32 arrays). Here they run through the same converters every entry point uses.
_SWIG_DEFAULT = -1


def _iarray(a, ndim, name):
    arr = _farray(a, ndim, name)       # the reference converts through its float typemap first, then truncates (swig/vector.i)
    return arr.astype(_np.int32)


def test_array(v):
    return _farray(v, 1, "v")


def test_vec_input(input):
    total = _np.float32(0)
    for v in _farray(input, 1, "input"):
        total = _np.float32(total + v)
    return float(total)


def test_ivec_input(input):
    return int(_iarray(input, 1, "input").sum())


def test_vec2_input(input):
    total = _np.float32(0)
    for v in _farray(input, 2, "input").ravel():
        total = _np.float32(total + v)
    return float(total)


def test_vec3_input(input):
    total = _np.float32(0)
    for v in _farray(input, 3, "input").ravel():
        total = _np.float32(total + v)
    return float(total)


def test_vec_output():
    return _np.full(3, _SWIG_DEFAULT, _np.float32)


def test_vec2_output():
    return _np.full((3, 3), _SWIG_DEFAULT, _np.float32)


def test_vec3_output():
    return _np.full((3, 3, 3), _SWIG_DEFAULT, _np.float32)


def test_ivec_output():
    return _np.full(3, _SWIG_DEFAULT, _np.int32)


def test_ivec2_output():
    return _np.full((3, 3), _SWIG_DEFAULT, _np.int32)


def test_ivec3_output():
    return _np.full((3, 3, 3), _SWIG_DEFAULT, _np.int32)


def test_vec_argout():
    return 0.0, _np.full(10, _SWIG_DEFAULT, _np.float32)


def test_vec2_argout():
    return 0.0, _np.full((10, 10), _SWIG_DEFAULT, _np.float32)


def test_not_implemented_exception():
    raise NotImplementedOnDevice("Function not yet implemented")


for _name in [n for n in list(globals()) if n.startswith("test_")]:
    globals()[_name].__test__ = False      # these are API functions, not tests for pytest to collect
del _name
