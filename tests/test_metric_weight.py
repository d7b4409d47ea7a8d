"""Curvilinear metric weight sqrt(det g) in scalar_product (orthogonal.py:268-276, tensorproductspace.py:376-379 of the
reference): 1-D spaces fold `system.sg / df` into the quadrature weights, a tensor product multiplies the samples by sg on the
mesh before the separable products (its factors keep sub-systems with sg = 1, coordinates.py:1254).  `forward` of a tensor
product does not see sg (tensorproductspace.py:395-417).  The `system` stand-in below follows the CoordSys protocol as far as the
transforms use it: `.sg` and `.base_scalars()`."""
import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O

r, th = sp.symbols("r theta", real=True)


class Polar:
    sg = r

    @staticmethod
    def base_scalars():
        return (r, th)


class Radial1D:
    sg = r**2

    @staticmethod
    def base_scalars():
        return (r,)


class Scaled:
    sg = sp.Integer(3)

    @staticmethod
    def base_scalars():
        return (r,)


def test_oracle_metric_weight_is_a_premultiplication():
    rng = np.random.default_rng(0)
    V = O.Legendre(12, domain=(0.5, 2.0))
    u = rng.standard_normal((5, 12))
    plain = V.scalar_product(u)
    V.system = Radial1D
    x = np.asarray(V.mesh())
    assert np.abs(V.scalar_product(u) - O.Legendre(12, domain=(0.5, 2.0)).scalar_product(u * x**2)).max() < 1e-13
    V.system = Scaled
    assert np.abs(V.scalar_product(u) - 3 * plain).max() < 1e-13
    T = O.TensorProductSpace(O.Legendre(10, domain=(0, 1)), O.Fourier(8), system=Polar)
    Tc = O.TensorProductSpace(O.Legendre(10, domain=(0, 1)), O.Fourier(8))
    w = rng.standard_normal((10, 8)) + 0j
    R = T.mesh()[0]
    assert np.abs(T.scalar_product(w) - Tc.scalar_product(w * R)).max() < 1e-13
    assert np.abs(T.forward(w) - Tc.forward(w)).max() == 0          # forward has no sg


@pytest.mark.gpu
def test_metric_weight_on_the_gpu_matches_the_oracle(cuda):
    import torch
    import jaxfun_b200 as jf
    rng = np.random.default_rng(1)

    def rel(a, b):
        return float(np.abs(a.cpu().numpy() - b).max() / np.abs(b).max())
    # 1-D: folded into the weights of the contraction table
    for sysm in (Radial1D, Scaled):
        V = jf.Legendre(24, domain=(0.5, 2.0), system=sysm)
        Vo = O.Legendre(24, domain=(0.5, 2.0))
        Vo.system = sysm
        u = rng.standard_normal((33, 24))
        assert rel(V.scalar_product(torch.from_numpy(u).to(cuda)), Vo.scalar_product(u)) < 1e-12
        assert rel(V.forward(torch.from_numpy(u).to(cuda)), Vo.forward(u)) < 1e-12
    # tensor product in polar coordinates: Legendre (radius) x Fourier (angle)
    T = jf.TensorProduct(jf.Legendre(32, domain=(0, 1)), jf.Fourier(16), system=Polar)
    To = O.TensorProductSpace(O.Legendre(32, domain=(0, 1)), O.Fourier(16), system=Polar)
    w = rng.standard_normal((32, 16)) + 1j * rng.standard_normal((32, 16))
    wd = torch.from_numpy(w).to(cuda)
    assert rel(T.scalar_product(wd), To.scalar_product(w)) < 1e-12
    assert rel(T.forward(wd), To.forward(w)) < 1e-12
    assert rel(T.backward(wd), To.backward(w)) < 1e-12
    # and a Cartesian product of the same spaces differs (the weight is really applied)
    Tc = jf.TensorProduct(jf.Legendre(32, domain=(0, 1)), jf.Fourier(16))
    assert rel(Tc.scalar_product(wd), To.scalar_product(w)) > 1e-3
