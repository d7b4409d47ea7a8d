"""Array type and tree helpers of the numpy stand-in for jax (golden-vector generation only)."""
import numpy as np


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _upd(self, fn):
        out = np.array(self.arr, copy=True)
        fn(out)
        return out.view(JArray)

    def set(self, v):
        def f(o):
            o[self.idx] = v
        return self._upd(f)

    def add(self, v):
        def f(o):
            o[self.idx] += v
        return self._upd(f)

    def multiply(self, v):
        def f(o):
            o[self.idx] *= v
        return self._upd(f)


class JArray(np.ndarray):
    """numpy array with jax's functional `.at[idx].set(v)` update syntax."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, JArray):
        return x.view(JArray)
    if isinstance(x, np.generic):
        return np.asarray(x).view(JArray)
    if isinstance(x, tuple):
        items = [wrap(v) for v in x]
        return type(x)(*items) if hasattr(x, "_fields") else tuple(items)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, (tuple, list)):
        items = [tree_map(f, *[t[i] for t in trees]) for i in range(len(t0))]
        return type(t0)(*items) if hasattr(t0, "_fields") else type(t0)(items)
    if isinstance(t0, dict):
        return {k: tree_map(f, *[t[k] for t in trees]) for k in t0}
    if t0 is None:
        return None
    return f(*trees)


def tree_leaves(t):
    if isinstance(t, (tuple, list)):
        out = []
        for v in t:
            out += tree_leaves(v)
        return out
    if isinstance(t, dict):
        out = []
        for v in t.values():
            out += tree_leaves(v)
        return out
    if t is None:
        return []
    return [t]
