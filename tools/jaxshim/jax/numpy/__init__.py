"""jax.numpy -> numpy (results viewed as JArray so `.at[]` works)."""
import numpy as _np

from .._core import JArray, wrap as _wrap
from . import fft  # noqa: F401

pi, e, inf, nan, newaxis = _np.pi, _np.e, _np.inf, _np.nan, _np.newaxis
ndarray = JArray


def _wrapped(fn):
    def f(*a, **k):
        return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def array(x, dtype=None, copy=True, **kw):
    return _np.array(x, dtype=dtype).view(JArray)


def asarray(x, dtype=None, **kw):
    return _np.asarray(x, dtype=dtype).view(JArray)


def __getattr__(name):
    obj = getattr(_np, name)
    if isinstance(obj, type) or not callable(obj):
        return obj
    return _wrapped(obj)
