"""Minimal numpy stand-in for the parts of the `jax` API the reference's transform path touches.

TEST INFRASTRUCTURE ONLY (tests/golden/make_golden.py): jax is not installable in this image, so
the golden vectors are produced by executing the reference's own source files
(/root/reference/src/jaxfun/galerkin/*.py, utils/fastgl.py, sharding.py ...) with numpy/scipy standing
in for jax.numpy / jax.scipy / jax.lax.  jit is the identity, vmap is a Python loop, scan/fori_loop
are Python loops, everything is float64/complex128 (jax_enable_x64).  Automatic differentiation is
not provided.
"""
import functools

import numpy as _np

from . import _core
from ._core import JArray as Array  # noqa: F401
from . import numpy, lax, scipy, typing, sharding, tree_util  # noqa: F401,E402


class _Config:
    def update(self, *a, **k):
        pass

    jax_enable_x64 = True


config = _Config()


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _slice(x, ax, i):
    if ax is None:
        return x
    return _core.tree_map(lambda a: _core.wrap(_np.take(_np.asarray(a), i, axis=ax)), x)


def vmap(fun, in_axes=0, out_axes=0, **kw):
    @functools.wraps(fun)
    def mapped(*args, **kwargs):
        ia = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(ia) == len(args), (ia, len(args))
        n = None
        for a, ax in zip(args, ia):
            if ax is not None:
                leaf = _core.tree_leaves(a)[0]
                n = _np.asarray(leaf).shape[ax]
                break
        assert n is not None, "vmap needs at least one mapped argument"
        outs = [fun(*[_slice(a, ax, i) for a, ax in zip(args, ia)], **kwargs) for i in range(n)]
        return _core.tree_map(lambda *o: _core.wrap(_np.stack([_np.asarray(v) for v in o], axis=out_axes)), *outs)
    return mapped


def _no_ad(*a, **k):
    raise NotImplementedError("the numpy stand-in for jax has no automatic differentiation")


grad = jacfwd = jacrev = hessian = value_and_grad = _no_ad


class _Device:
    platform = "cpu"
    id = 0

    def __repr__(self):
        return "CpuDevice(id=0)"


def devices(*a):
    return [_Device()]


local_devices = devices


def local_device_count(*a):
    return 1


device_count = local_device_count


def process_index():
    return 0


def device_put(x, *a, **k):
    return x


def shard_map(f, **kw):
    raise NotImplementedError("shard_map is not emulated")


def make_array_from_single_device_arrays(*a, **k):
    raise NotImplementedError
