"""jax.lax control flow as Python loops."""
import numpy as _np

from ._core import tree_leaves, tree_map, wrap as _wrap


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
    else:
        n = _np.asarray(tree_leaves(xs)[0]).shape[0]
    idx = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in idx:
        x = None if xs is None else tree_map(lambda a: _wrap(_np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys[i] = y
    if n == 0 or ys[0] is None:
        return carry, None
    return carry, tree_map(lambda *o: _wrap(_np.stack([_np.asarray(v) for v in o])), *ys)


def fori_loop(lower, upper, body, init, **kw):
    v = init
    for i in range(int(lower), int(upper)):
        v = body(i, v)
    return v


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def select(pred, a, b):
    return _wrap(_np.where(pred, a, b))


def dynamic_slice(x, start, sizes):
    sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, sizes))
    return _wrap(_np.asarray(x)[sl])


def exp(x):
    return _wrap(_np.exp(x))


def all_to_all(*a, **k):
    raise NotImplementedError


psum = all_to_all
