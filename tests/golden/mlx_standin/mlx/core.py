"""Minimal NumPy-backed stand-in for the `mlx.core` API surface that the reference's hot-path
modules use -- TEST INFRASTRUCTURE ONLY, used by tests/golden/make_reference_golden.py.

Why it exists: the reference (service/optimized_vector_store.py, performance/mlx_optimized.py)
is pure Python over `mlx.core`, and `mlx` (pinned `mlx>=0.25.2`, requirements.txt:7) has no
wheel in this image and there is no network.  With this module on sys.path the reference's
OWN Python files import and run unmodified from /root/reference, so the golden fixtures pin
everything the reference itself decides -- op order, clamps, reshapes, slicing, the filter and
id mapping logic, empty-store / k > N / bad-shape behaviour, return types -- and leave only the
arithmetic inside the mlx primitives to this stand-in:

  * every primitive computes in IEEE fp32 with NumPy (matmul = the BLAS sgemm NumPy links);
  * `argsort` is a stable ascending sort (ties -> lower index first), the documented assumption
    about mlx's CPU argsort;
  * `compile` is the identity decorator, `eval` a no-op (mlx is lazy, NumPy is eager).

Anything mlx does differently from NumPy at the last ulp (accumulation order inside matmul,
fused multiply-adds) is outside what these fixtures can pin; BASELINE.json's tolerance
(scores within 1e-5, ids exact outside 1e-6 ties) is far wider than that.
"""
from __future__ import annotations

import numpy as _np

float32 = _np.float32
float16 = _np.float16
int32 = _np.int32
uint32 = _np.uint32
int64 = _np.int64


class array(_np.ndarray):
    """`mx.array(obj, dtype=...)`: an ndarray subclass, so `isinstance(x, mx.array)`, `.ndim`,
    `.shape`, `.T`, `.reshape`, `.flatten`, `.astype`, `.tolist`, `.size`, slicing and fancy
    indexing behave as the reference expects, and results of arithmetic stay `mx.array`."""

    def __new__(cls, obj, dtype=None):
        if isinstance(obj, _np.ndarray) and dtype is None:
            dtype = obj.dtype
        if dtype is None:
            probe = _np.asarray(obj)
            # mlx defaults: Python floats -> float32, ints -> int32
            dtype = float32 if probe.dtype.kind == "f" else (int32 if probe.dtype.kind in "iu" else probe.dtype)
        return _np.array(obj, dtype=dtype).view(cls)


def _wrap(x):
    return x.view(array) if isinstance(x, _np.ndarray) else array(x)


def compile(fn=None, **_kw):  # noqa: A001 - mirrors mx.compile
    if fn is None:
        return lambda f: f
    return fn


def eval(*_args, **_kw):  # noqa: A001 - mirrors mx.eval (NumPy is eager)
    return None


class _Linalg:
    @staticmethod
    def norm(x, axis=None, keepdims=False):
        x = _np.asarray(x)
        return _wrap(_np.sqrt(_np.sum(_np.square(x), axis=axis, keepdims=keepdims, dtype=x.dtype)))


linalg = _Linalg()


def maximum(a, b):
    return _wrap(_np.maximum(a, b))


def matmul(a, b):
    return _wrap(_np.matmul(_np.asarray(a), _np.asarray(b)))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    x = _np.asarray(x)
    return _wrap(_np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype))


def sqrt(x):
    return _wrap(_np.sqrt(_np.asarray(x)))


def square(x):
    return _wrap(_np.square(_np.asarray(x)))


def argsort(x, axis=-1):
    return _wrap(_np.argsort(_np.asarray(x), axis=axis, kind="stable").astype(uint32))


def concatenate(arrays, axis=0):
    return _wrap(_np.concatenate([_np.asarray(a) for a in arrays], axis=axis))


def stack(arrays, axis=0):
    return _wrap(_np.stack([_np.asarray(a) for a in arrays], axis=axis))


def zeros(shape, dtype=float32):
    return _wrap(_np.zeros(shape, dtype=dtype))


def savez(path, **arrays):
    _np.savez(path, **{k: _np.asarray(v) for k, v in arrays.items()})


def load(path):
    with _np.load(path) as z:
        return {k: _wrap(z[k]) for k in z.files}


class _Random:
    @staticmethod
    def normal(shape=(), dtype=float32, **_kw):
        return _wrap(_np.random.standard_normal(shape).astype(dtype))


random = _Random()
