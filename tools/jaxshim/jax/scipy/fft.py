import scipy.fft as _f

from .._core import wrap as _wrap


def dct(x, type=2, n=None, axis=-1, norm=None):
    return _wrap(_f.dct(x, type=type, n=n, axis=axis, norm=norm))


def idct(x, type=2, n=None, axis=-1, norm=None):
    return _wrap(_f.idct(x, type=type, n=n, axis=axis, norm=norm))
