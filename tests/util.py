"""Helpers shared by the tests: golden fixtures, structure descriptors, the parity metric."""
import os

import numpy as np

from oracle import bindings as B

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def parse_spec(spec):
    """Fixture structure specs are stored as repr((kind, args, kwargs)) strings of plain literals."""
    import ast
    return ast.literal_eval(str(np.asarray(spec).ravel()[0]))


def oracle_structure(spec):
    kind, args, kw = spec
    if kind == "single":
        return B.make_structure(*args, **kw)
    if kind == "cv":
        return B.cross_validation(oracle_structure(args[0]), args[1])
    if kind == "multiple":
        return B.multiple_structure(*[oracle_structure(a) for a in args])
    raise ValueError(kind)


_CLASSES = ("BarnesStructure", "CressmanStructure", "SoarStructure", "ToarStructure", "PowerlawStructure", "LinearStructure")


def product_structure(gpp, spec):
    kind, args, kw = spec
    if kind == "single":
        cls = getattr(gpp, _CLASSES[args[0]])
        if args[0] == B.CRESSMAN:
            return cls(*args[1:4])
        return cls(*args[1:])
    if kind == "cv":
        return gpp.CrossValidation(product_structure(gpp, args[0]), args[1])
    if kind == "multiple":
        return gpp.MultipleStructure(*[product_structure(gpp, a) for a in args])
    raise ValueError(kind)


def assert_same_nan(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    mism = np.isnan(a) != np.isnan(b)
    assert not mism.any(), "NaN pattern differs at %d of %d cells, first %s" % (mism.sum(), a.size, np.argwhere(mism)[:3].tolist())


def assert_bit_exact(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        assert_same_nan(a, b)
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    else:
        bad = a != b
    assert not bad.any(), "%s: %d of %d values differ, first at %s: %r vs %r" % (
        what, bad.sum(), a.size, np.argwhere(bad)[:3].tolist(), a[bad][:3], b[bad][:3])


def assert_close(got, want, scale, rtol=1e-5, what="", allow_outliers=0):
    """SURVEY.md section 8(d) parity metric: |got - want| <= rtol * max(|want|, scale). Returns the worst ratio."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert_same_nan(got, want)
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), scale)
    n_bad = int((err > rtol).sum())
    worst = float(err.max()) if err.size else 0.0
    assert n_bad <= allow_outliers, "%s: %d values beyond rtol=%g (worst %.3e)" % (what, n_bad, rtol, worst)
    return worst
