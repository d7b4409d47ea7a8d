import scipy.linalg as _l

from .._core import wrap as _wrap


def __getattr__(name):
    fn = getattr(_l, name)
    return lambda *a, **k: _wrap(fn(*a, **k))
