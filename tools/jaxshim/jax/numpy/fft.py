import numpy as _np

from .._core import wrap as _wrap


def __getattr__(name):
    fn = getattr(_np.fft, name)
    return lambda *a, **k: _wrap(fn(*a, **k))
