"""`import drjit as dr`: the array functions the reference calls, on numpy.
Host-side glue only -- the acquisition itself never runs through these."""
import math as _math

import numpy as _np

pi = _math.pi
inf = _math.inf


def _a(x):
    return _np.asarray(x.numpy() if hasattr(x, "numpy") and not isinstance(x, _np.ndarray) else x)


def _like(ref, out):
    t = type(ref)
    try:
        return out.view(t) if isinstance(ref, _np.ndarray) and isinstance(out, _np.ndarray) and out.dtype == _np.float32 else out
    except Exception:
        return out


def dot(a, b): return _np.sum(_a(a) * _a(b), axis=-1)
def norm(a): return _np.sqrt(_np.sum(_a(a) * _a(a), axis=-1))
def normalize(a):
    v = _a(a)
    return _like(a, (v / _np.sqrt(_np.sum(v * v, axis=-1, keepdims=True))).astype(v.dtype, copy=False))
def cross(a, b): return _like(a, _np.cross(_a(a), _a(b)).astype(_np.float32))
def select(m, a, b): return _np.where(_a(m), _a(a), _a(b))
def deg2rad(x): return _a(x) * (_math.pi / 180.0) if not isinstance(x, (int, float)) else x * _math.pi / 180.0
def rad2deg(x): return _a(x) * (180.0 / _math.pi) if not isinstance(x, (int, float)) else x * 180.0 / _math.pi
def sin(x): return _np.sin(_a(x))
def cos(x): return _np.cos(_a(x))
def acos(x): return _np.arccos(_a(x))
def exp(x): return _np.exp(_a(x))
def sqrt(x): return _np.sqrt(_a(x))
def rsqrt(x): return 1.0 / _np.sqrt(_a(x))
def abs(x): return _np.abs(_a(x))
def square(x): return _a(x) * _a(x)
def maximum(a, b): return _np.maximum(_a(a), _a(b))
def minimum(a, b): return _np.minimum(_a(a), _a(b))
def clamp(x, lo, hi): return _np.clip(_a(x), lo, hi)
clip = clamp
def floor(x): return _np.floor(_a(x))
def round(x): return _np.rint(_a(x))            # round-half-to-even, as dr.round
def mean(x): return _np.mean(_a(x))
def max(x): return _np.max(_a(x))
def min(x): return _np.min(_a(x))
def sum(x): return _np.sum(_a(x))
def isinf(x): return _np.isinf(_a(x))
def isnan(x): return _np.isnan(_a(x))
def detach(x): return x
def eval(*a): return None
def backward(*a): raise NotImplementedError("automatic differentiation is outside the hot path")
def array(x): return _np.asarray(x)
def print(*a, **k): __builtins__["print"](*a, **k) if isinstance(__builtins__, dict) else __builtins__.print(*a, **k)


def _mk(dtype, arr):
    if isinstance(dtype, type) and issubclass(dtype, _np.ndarray):
        return dtype(arr)
    return _np.asarray(arr, dtype=_np.float32)


def zeros(dtype, shape=1): return _mk(dtype, _np.zeros(shape, dtype=_np.float32))
def ones(dtype, shape=1): return _mk(dtype, _np.ones(shape, dtype=_np.float32))
def full(dtype, value, shape=1): return _mk(dtype, _np.full(shape, value, dtype=_np.float32))
def arange(dtype, *args): return _mk(dtype, _np.arange(*args, dtype=_np.float32))
def linspace(dtype, start, stop, num, endpoint=True): return _mk(dtype, _np.linspace(start, stop, num, endpoint=endpoint, dtype=_np.float32))


def gather(dtype, source, index, active=True): return _mk(dtype, _a(source)[_a(index).astype(_np.int64)])


def scatter(target, value, index, active=True):
    _np.asarray(target)[_a(index).astype(_np.int64)] = _a(value)


class ReduceOp:
    Add = "add"


def scatter_reduce(op, target, value, index, active=True):
    _np.add.at(_np.asarray(target), _a(index).astype(_np.int64)[_a(active)] if not isinstance(active, bool) else _a(index).astype(_np.int64),
               _a(value))


def while_loop(state, cond, body, **kw):
    while bool(_np.all(_a(cond(*state)))):
        state = body(*state)
    return state


class _LLVM:
    Float = _np.float32


llvm = _LLVM()
