"""`symnp`: the slice of the `jax.numpy` namespace the reference's dynamics are written in, on numpy OBJECT
arrays of sympy expressions.

The reference defines f(x,u), l(x,u) with `jax.numpy` (imported as `np`) and lets jax autodiff trace them
(flatquad_landing_experiment.py:277-329, experiment_orbits.py:24-40, experiment_lqr_statepenalty.py:23-38).
codegen.py traces the same function bodies symbolically instead: called with object arrays of sympy symbols and
with this module standing in for `np`, they return sympy expressions that are differentiated and printed as CUDA.
Numeric arguments fall through to numpy, so the same f, l also evaluate on floats.
"""
from __future__ import annotations

import numpy as _np
import sympy as _sp

pi = _np.pi
inf = _np.inf
newaxis = None
float32, float64 = _np.float32, _np.float64


def _symbolic(*args) -> bool:
    for a in args:
        if isinstance(a, _sp.Basic):
            return True
        if isinstance(a, _np.ndarray):
            if a.dtype == object and any(isinstance(e, _sp.Basic) for e in a.view(_np.ndarray).ravel().tolist()):
                return True
        elif isinstance(a, (list, tuple)) and _symbolic(*a):
            return True
    return False


class SymArray(_np.ndarray):
    """Object array of sympy expressions that behaves like a jax array where numpy differs: indexing and
    arithmetic keep 0-d arrays (so `x[0] ** 2` still has .reshape / .squeeze / .T, experiment_orbits.py:25),
    and reshape() without arguments gives a scalar (experiment_orbits.py:40)."""

    def __new__(cls, obj):
        return _np.asarray(obj, dtype=object).view(cls)

    def __getitem__(self, idx):
        r = super().__getitem__(idx)
        return r if isinstance(r, _np.ndarray) else _np.asarray(r, dtype=object).view(SymArray)

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        return _np.asarray(arr).view(SymArray)

    def reshape(self, *shape, **kw):
        return super().reshape(*(shape if shape else ((),)), **kw)


def array(obj, dtype=None):
    if _symbolic(obj):
        return _np.array(obj, dtype=object)
    return _np.array(obj, dtype=float if dtype is None else dtype)


asarray = array


def _unary(sym_fn, np_fn):
    def fn(a):
        if _symbolic(a):
            if isinstance(a, _np.ndarray):
                src = a.view(_np.ndarray)
                out = _np.empty(a.shape, dtype=object)
                for idx in _np.ndindex(a.shape):
                    out[idx] = sym_fn(src[idx])
                return out.view(type(a))
            return sym_fn(a)
        return np_fn(a)
    fn.__name__ = np_fn.__name__
    return fn


sin = _unary(_sp.sin, _np.sin)
cos = _unary(_sp.cos, _np.cos)
tan = _unary(_sp.tan, _np.tan)
tanh = _unary(_sp.tanh, _np.tanh)
exp = _unary(_sp.exp, _np.exp)
log = _unary(_sp.log, _np.log)
sqrt = _unary(_sp.sqrt, _np.sqrt)
abs = absolute = _unary(_sp.Abs, _np.abs)  # noqa: A001
square = _unary(lambda e: e ** 2, _np.square)
deg2rad, rad2deg = _np.deg2rad, _np.rad2deg


def _binary(sym_fn, np_fn):
    def fn(a, b):
        if _symbolic(a, b):
            a, b = _np.broadcast_arrays(_np.asarray(a, dtype=object).view(_np.ndarray),
                                        _np.asarray(b, dtype=object).view(_np.ndarray))
            out = _np.empty(a.shape, dtype=object)
            for idx in _np.ndindex(a.shape):
                out[idx] = sym_fn(_sp.sympify(a[idx]), _sp.sympify(b[idx]))
            return out if out.shape else out.item()
        return np_fn(a, b)
    return fn


maximum = _binary(_sp.Max, _np.maximum)
minimum = _binary(_sp.Min, _np.minimum)
arctan2 = _binary(_sp.atan2, _np.arctan2)
power = _binary(lambda a, b: a ** b, _np.power)


def clip(a, lo, hi):
    return minimum(maximum(a, lo), hi)


def zeros(shape, dtype=None):
    return _np.zeros(shape)


def ones(shape, dtype=None):
    return _np.ones(shape)


eye, diag, arange, linspace = _np.eye, _np.diag, _np.arange, _np.linspace


def _objs(seq):
    return [_np.asarray(a, dtype=object) if _symbolic(a) else _np.asarray(a) for a in seq]


def concatenate(seq, axis=0):
    return _np.concatenate(_objs(seq), axis=axis)


def stack(seq, axis=0):
    return _np.stack(_objs(seq), axis=axis)


def hstack(seq):
    return _np.hstack(_objs(seq))


def vstack(seq):
    return _np.vstack(_objs(seq))


def dot(a, b):
    return _np.dot(_np.asarray(a, dtype=object) if _symbolic(a) else a, _np.asarray(b, dtype=object) if _symbolic(b) else b)


def sum(a, axis=None):  # noqa: A001
    return _np.sum(a, axis=axis)


def squeeze(a, axis=None):
    return _np.squeeze(a, axis=axis)


def reshape(a, shape):
    return _np.reshape(a, shape)


class linalg:  # noqa: N801  (namespace, as in numpy)
    @staticmethod
    def norm(a):
        a = _np.asarray(a, dtype=object) if _symbolic(a) else _np.asarray(a)
        if a.dtype == object:
            return _sp.sqrt(_np.sum(a.ravel() ** 2))
        return _np.linalg.norm(a)
