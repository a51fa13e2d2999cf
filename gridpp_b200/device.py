"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out, no host copies, no sync.

torch is used only as the owner of device memory and streams; every call forwards raw pointers and the
current CUDA stream to the ``*_device`` functions of the C ABI (include/gridpp_b200.h). These are the calls
bench.py times for the kernel-only (`value`) number and the ones a multi-GPU driver shards by rows.
"""
import ctypes as _C

import numpy as _np
import torch

from . import _lib
from ._lib import check as _check, lib as _libc
from . import Grid, Points, _farray, _fptr


def _stream_ptr():
    return _C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError("expected a contiguous float32 CUDA tensor")
    return _C.c_void_p(t.data_ptr())


class ObservationState:
    """Observation side of an OI call, resident on the device (gpp_oi_obs): the bucket grid over the valid
    observations with their innovations and variance ratios. Build once, reuse for every row block / call."""

    def __init__(self, points, pobs, obs_variance, pbackground, structure, bvariance_at_points=None):
        if not isinstance(points, Points):
            raise ValueError("points must be a Points object")
        n = points.size()
        obs, ovar, pbg = _farray(pobs, 1, "pobs"), _farray(obs_variance, 1, "obs_variance"), _farray(pbackground, 1, "pbackground")
        if obs.size != n or ovar.size != n or pbg.size != n:
            raise ValueError("Observations / variances / background and points size mismatch")
        pbvar = _farray(bvariance_at_points, 1, "bvariance_at_points") if bvariance_at_points is not None else None
        self.points = points
        self.structure = structure
        self._handle = _C.c_void_p()
        _check(_libc.gpp_oi_obs_create(points._set._handle, _fptr(obs), _fptr(ovar), _fptr(pbg), _fptr(pbvar),
                                       _C.byref(structure._desc), _C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h and _libc is not None:
            _libc.gpp_oi_obs_destroy(h)
            self._handle = None


def optimal_interpolation(bgrid, background, obs_state, max_points, allow_extrapolation=True, out=None, bvariance=None,
                          out_variance=None, first=0, count=None):
    """Analyses background points [first, first+count) of `bgrid` (a Grid or Points; flattened row-major).
    `background`, `out`, `bvariance`, `out_variance` are float32 CUDA tensors covering the WHOLE field, indexed like
    the flattened grid; only the requested range is read / written. Asynchronous on the current stream."""
    n = bgrid._set.n
    if background.numel() != n:
        raise ValueError("background has %d elements, the grid %d" % (background.numel(), n))
    if out is None:
        out = torch.empty_like(background)
    if count is None:
        count = n - first
    _check(_libc.gpp_optimal_interpolation_device(bgrid._set._handle, int(first), int(count), _ptr(background), _ptr(bvariance),
                                                  obs_state._handle, _C.byref(obs_state.structure._desc), int(max_points),
                                                  int(bool(allow_extrapolation)), _ptr(out), _ptr(out_variance), _stream_ptr()))
    return out


class Workspace:
    """Explicit launch workspace of the register OI path (gpp_oi_workspace_bytes): zero-filled once, then owned by
    one launch in flight at a time. With it gpp_optimal_interpolation_device_ws performs no allocation at all."""

    def __init__(self, device=None):
        self.bytes = int(_libc.gpp_oi_workspace_bytes())
        self.buf = torch.zeros(max(self.bytes, 1), dtype=torch.uint8, device=device or "cuda")


def optimal_interpolation_ws(bgrid, background, obs_state, max_points, workspace, allow_extrapolation=True, out=None, bvariance=None,
                             out_variance=None, first=0, count=None):
    """optimal_interpolation with a caller-owned Workspace (one per launch in flight)."""
    n = bgrid._set.n
    if background.numel() != n:
        raise ValueError("background has %d elements, the grid %d" % (background.numel(), n))
    if out is None:
        out = torch.empty_like(background)
    if count is None:
        count = n - first
    _check(_libc.gpp_optimal_interpolation_device_ws(bgrid._set._handle, int(first), int(count), _ptr(background), _ptr(bvariance),
                                                     obs_state._handle, _C.byref(obs_state.structure._desc), int(max_points),
                                                     int(bool(allow_extrapolation)), _ptr(out), _ptr(out_variance),
                                                     _C.c_void_p(workspace.buf.data_ptr()), workspace.bytes, _stream_ptr()))
    return out


class EnsembleObservationState:
    """Observation side of optimal_interpolation_ensi, resident on the device (gpp_ensi_obs). `member_valid`: per-member
    flags (False = the member has an invalid value somewhere in the background and is left untouched,
    oi_ensi.cpp:187-201); None = all valid; `valid_members(background)` computes them from a device-resident field."""

    def __init__(self, points, pobs, psigmas, pbackground, structure, member_valid=None):
        if not isinstance(points, Points):
            raise ValueError("points must be a Points object")
        n = points.size()
        obs, sig = _farray(pobs, 1, "pobs"), _farray(psigmas, 1, "psigmas")
        pbg = _farray(pbackground, 2, "pbackground")
        if obs.size != n or sig.size != n or pbg.shape[0] != n:
            raise ValueError("Observations / sigmas / background and points size mismatch")
        self.nE = int(pbg.shape[1])
        flags = None
        if member_valid is not None:
            flags = _np.ascontiguousarray(_np.asarray(member_valid).astype(_np.int32))
            if flags.size != self.nE:
                raise ValueError("member_valid must have one flag per ensemble member")
        self.points = points
        self.structure = structure
        self._handle = _C.c_void_p()
        _check(_libc.gpp_ensi_obs_create(points._set._handle, _fptr(obs), _fptr(sig), _fptr(pbg), self.nE,
                                         flags.ctypes.data_as(_lib.ip) if flags is not None else None, _C.byref(structure._desc),
                                         _C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h and _libc is not None:
            _libc.gpp_ensi_obs_destroy(h)
            self._handle = None


def valid_members(background):
    """Per-member flags of a device-resident (..., E) ensemble: False where the member has an invalid value anywhere."""
    nE = int(background.shape[-1])
    flags = _np.zeros(max(nE, 1), _np.int32)
    _check(_libc.gpp_ensi_valid_members_device(_ptr(background), background.numel() // max(nE, 1), nE, flags.ctypes.data_as(_lib.ip),
                                               _stream_ptr()))
    return flags[:nE].astype(bool)


def optimal_interpolation_ensi(bgrid, background, obs_state, max_points, allow_extrapolation=True, out=None, num_skipped=None,
                               first=0, count=None):
    """Analyses background points [first, first+count) of `bgrid`; `background` / `out` are float32 CUDA tensors of the WHOLE
    (points, E) ensemble (they may be the same tensor). One kernel launch on the current stream, no synchronisation.
    `num_skipped`: optional int32 CUDA tensor (1 element) counting the points skipped for a numerically bad Pinv."""
    n = bgrid._set.n
    nE = obs_state.nE
    if background.numel() != n * nE:
        raise ValueError("background has %d elements, the grid x members %d" % (background.numel(), n * nE))
    if out is None:
        out = torch.empty_like(background)
    if count is None:
        count = n - first
    skipped = None
    if num_skipped is not None:
        if not (num_skipped.is_cuda and num_skipped.dtype == torch.int32):
            raise ValueError("num_skipped must be an int32 CUDA tensor")
        skipped = _C.c_void_p(num_skipped.data_ptr())
    _check(_libc.gpp_optimal_interpolation_ensi_device(bgrid._set._handle, int(first), int(count), _ptr(background), nE, obs_state._handle,
                                                       _C.byref(obs_state.structure._desc), int(max_points), int(bool(allow_extrapolation)),
                                                       _ptr(out), skipped, _stream_ptr()))
    return out


def neighbourhood(field, halfwidth, statistic, out=None, row0=0, n_rows_out=None):
    """field: (rows, nx) float32 CUDA tensor (a tile plus its halo rows); computes output rows
    [row0, row0 + n_rows_out) into `out` (n_rows_out, nx)."""
    rows, nx = field.shape
    if n_rows_out is None:
        n_rows_out = rows - row0
    if out is None:
        out = torch.empty((n_rows_out, nx), dtype=torch.float32, device=field.device)
    _check(_libc.gpp_neighbourhood_device(_ptr(field), rows, nx, int(row0), int(n_rows_out), int(halfwidth), int(statistic),
                                          _ptr(out), _stream_ptr()))
    return out


def neighbourhood_quantile_fast(field, quantile, halfwidth, thresholds, out=None, row0=0, n_rows_out=None):
    rows, nx = field.shape
    if n_rows_out is None:
        n_rows_out = rows - row0
    if out is None:
        out = torch.empty((n_rows_out, nx), dtype=torch.float32, device=field.device)
    thr = _farray(thresholds, 1, "thresholds")
    qf = None
    q = float("nan")
    if torch.is_tensor(quantile):
        qf = quantile
        if qf.shape != field.shape:
            raise ValueError("the quantile field must have the shape of the input tile")
    else:
        q = float(quantile)
    _check(_libc.gpp_neighbourhood_quantile_fast_device(_ptr(field), rows, nx, int(row0), int(n_rows_out), q, _ptr(qf),
                                                        int(halfwidth), _fptr(thr), thr.size, _ptr(out), _stream_ptr()))
    return out
