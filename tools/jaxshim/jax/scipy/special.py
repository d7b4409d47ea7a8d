import scipy.special as _s

from .._core import wrap as _wrap


def __getattr__(name):
    fn = getattr(_s, name)
    return lambda *a, **k: _wrap(fn(*a, **k))
