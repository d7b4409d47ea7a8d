"""FFT / DCT fast kernels at every supported length, contiguous and strided axes, real and complex
data, with padding (backward) and truncation (forward), against the CPU oracle (rel 1e-12)."""
import numpy as np
import pytest
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
from jaxfun_b200.engine import fast_path_available

pytestmark = pytest.mark.gpu
SIZES = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 48, 96, 192, 384, 80, 160, 320]


def relerr(a, b):
    a = a.detach().cpu().numpy()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / np.abs(b).max())


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.parametrize("n", SIZES)
def test_fast_path_is_reported(cuda, n):
    assert fast_path_available(L.BASIS_CHEBYSHEV, n, L.F64)
    assert fast_path_available(L.BASIS_CHEBYSHEV, n, L.C128)
    assert fast_path_available(L.BASIS_FOURIER, n, L.C128)
    assert not fast_path_available(L.BASIS_FOURIER, n, L.F64)
    assert not fast_path_available(L.BASIS_CHEBYSHEV, n + 2, L.F64)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("layout", ["last", "middle", "first"])
def test_fourier_fft(cuda, n, layout):
    rng = np.random.default_rng(n)
    N = n                      # unpadded
    shape, axis = {"last": ((7, N), 1), "middle": ((3, N, 10), 1), "first": ((N, 9), 0)}[layout]
    o, p = O.Fourier(N, domain=(0.0, 1.0)), jf.Fourier(N, domain=(0.0, 1.0))
    c = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    u_ref = o.backward(c, axis=axis)
    assert relerr(p.backward(dev(c, cuda), axis=axis), u_ref) < 1e-12
    assert relerr(p.forward(dev(u_ref, cuda), axis=axis), o.forward(u_ref, axis=axis)) < 1e-12
    assert relerr(p.scalar_product(dev(u_ref, cuda), axis=axis), o.scalar_product(u_ref, axis=axis)) < 1e-12
    for k in (1, 2, 3):
        ref = o.backward_primitive(c, k=k, axis=axis)
        assert relerr(p.backward_primitive(dev(c, cuda), k=k, axis=axis), ref) < 1e-12


@pytest.mark.parametrize("n", SIZES)
def test_fourier_padding_truncation(cuda, n):
    """3/2-rule style padding: N = 2n/3 (even) modes on n points (tests/galerkin/test_fourier.py:7-16)."""
    rng = np.random.default_rng(n + 1)
    N = (2 * n // 3) // 2 * 2
    o, p = O.Fourier(N), jf.Fourier(N)
    c = rng.standard_normal((5, N)) + 1j * rng.standard_normal((5, N))
    u_ref = o.backward(c, N=n)
    u = p.backward(dev(c, cuda), N=n)
    assert tuple(u.shape) == (5, n) and relerr(u, u_ref) < 1e-12
    assert relerr(p.forward(u), c) < 1e-12
    assert relerr(p.scalar_product(dev(u_ref, cuda)), o.scalar_product(u_ref)) < 1e-12
    assert relerr(p.backward_primitive(dev(c, cuda), k=1, N=n), o.backward_primitive(c, k=1, N=n)) < 1e-12


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("layout", ["last", "last_odd_lines", "middle", "first"])
@pytest.mark.parametrize("cplx", [False, True])
def test_chebyshev_dct(cuda, n, layout, cplx):
    rng = np.random.default_rng(n + 7)
    shape, axis = {"last": ((6, n), 1), "last_odd_lines": ((5, n), 1), "middle": ((3, n, 10), 1),
                   "first": ((n, 8), 0)}[layout]
    o, p = O.Chebyshev(n, domain=(-2.0, 3.0)), jf.Chebyshev(n, domain=(-2.0, 3.0))
    c = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
    u_ref = o.backward(c, axis=axis)
    assert relerr(p.backward(dev(c, cuda), axis=axis), u_ref) < 1e-12
    assert relerr(p.forward(dev(u_ref, cuda), axis=axis), o.forward(u_ref, axis=axis)) < 1e-12
    assert relerr(p.scalar_product(dev(u_ref, cuda), axis=axis), o.scalar_product(u_ref, axis=axis)) < 1e-12


@pytest.mark.parametrize("n", SIZES)
def test_chebyshev_padding_truncation(cuda, n):
    rng = np.random.default_rng(n + 3)
    N = 2 * n // 3
    o, p = O.Chebyshev(N), jf.Chebyshev(N)
    c = rng.standard_normal((4, N))
    u_ref = o.backward(c, N=n)
    u = p.backward(dev(c, cuda), N=n)
    assert tuple(u.shape) == (4, n) and relerr(u, u_ref) < 1e-12
    assert relerr(p.forward(u), c) < 1e-12
    assert relerr(p.scalar_product(dev(u_ref, cuda)), o.scalar_product(u_ref)) < 1e-12


def test_chebyshev_odd_inner_falls_back_to_table(cuda):
    rng = np.random.default_rng(1)
    o, p = O.Chebyshev(64), jf.Chebyshev(64)
    c = rng.standard_normal((64, 7))           # real data, odd inner extent: no DCT tile
    assert relerr(p.backward(dev(c, cuda), axis=0), o.backward(c, axis=0)) < 1e-12


def test_float32_fast_kernels(cuda):
    rng = np.random.default_rng(8)
    for n in (64, 256, 1024):
        o, p = O.Chebyshev(n), jf.Chebyshev(n)
        c = rng.standard_normal((6, n)).astype(np.float32)
        assert relerr(p.backward(dev(c, cuda)), o.backward(c.astype(np.float64)).astype(np.float32)) < 1e-5
        o, p = O.Fourier(n), jf.Fourier(n)
        c = (rng.standard_normal((6, n)) + 1j * rng.standard_normal((6, n))).astype(np.complex64)
        assert relerr(p.backward(dev(c, cuda)), o.backward(c.astype(np.complex128)).astype(np.complex64)) < 1e-5


def test_batched_1d_config_c3_properties(cuda):
    """BASELINE configs[2] shape [65536, 1024]: round trip + Parseval at full size (no oracle run)."""
    torch.manual_seed(3)
    n = 1024
    F = jf.Fourier(n)
    c = torch.randn(65536, n, dtype=torch.complex128, device=cuda)
    u = F.backward(c)
    # Parseval: sum |u_j|^2 = n sum |c_k|^2 for the unnormalised inverse DFT
    lhs = float((u.abs() ** 2).sum())
    rhs = float(n * (c.abs() ** 2).sum())
    assert abs(lhs - rhs) < 1e-12 * rhs
    back = F.forward(u)
    assert float((back - c).abs().max()) < 1e-12 * float(c.abs().max())


@pytest.mark.parametrize("names,N", [
    (("Chebyshev", "Chebyshev", "Chebyshev"), (16, 128, 128)),
    (("Chebyshev", "Chebyshev", "Chebyshev"), (8, 256, 256)),
    (("Fourier", "Fourier", "Fourier"), (16, 128, 128)),
    (("Legendre", "Chebyshev", "Chebyshev"), (24, 128, 128)),
    (("Fourier", "Chebyshev", "Chebyshev"), (8, 128, 128)),
])
def test_plane_fused_pair_vs_oracle(cuda, names, N, monkeypatch):
    """Shapes whose last two axes run as ONE plane-fused launch (kernels_fft2_pair.cu: ticketed A/B tiles,
    ring-buffered intermediate; opt-in with JFX_PAIR=1 at plan creation) against the oracle, all three
    directions."""
    monkeypatch.setenv("JFX_PAIR", "1")
    import jaxfun_oracle as O
    import jaxfun_b200 as jf
    rng = np.random.default_rng(sum(N))
    To = O.TensorProductSpace(*[getattr(O, n)(Ni) for n, Ni in zip(names, N)])
    Tp = jf.TensorProduct(*[getattr(jf, n)(Ni) for n, Ni in zip(names, N)])
    cplx = "Fourier" in names
    c = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0)
    u_ref = To.backward(c)
    cd = torch.from_numpy(c).to(cuda)
    plan = Tp._plan(2, cd)   # OP_BACKWARD
    assert plan.launches == 2, "expected axis-0 pass + one plane-fused launch"
    u = Tp.backward(cd)
    scale = np.abs(u_ref).max()
    assert np.abs(u.cpu().numpy() - u_ref).max() < 1e-12 * scale
    ud = torch.from_numpy(u_ref).to(cuda)
    f_ref, s_ref = To.forward(u_ref), To.scalar_product(u_ref)
    assert np.abs(Tp.forward(ud).cpu().numpy() - f_ref).max() < 1e-12 * np.abs(f_ref).max()
    assert np.abs(Tp.scalar_product(ud).cpu().numpy() - s_ref).max() < 1e-12 * np.abs(s_ref).max()
    # repeated execution reuses ring + counters
    for _ in range(3):
        u2 = Tp.backward(cd)
    assert torch.equal(u2, u)


@pytest.mark.parametrize("kind,shape,axis", [
    # shapes with >= 2 tiles per resident CTA, the envelope in which the streaming kernel engages
    ("Chebyshev", (16384, 256), 1), ("Chebyshev", (256, 16384), 0), ("Chebyshev", (64, 128, 1024), 1),
    ("Fourier", (2048, 1024), 1), ("Fourier", (128, 32768), 0),
])
def test_streaming_variant_vs_oracle(cuda, kind, shape, axis, monkeypatch):
    """kernels_fft2_stream.cu (persistent CTAs + bulk-async prefetch; opt-in with JFX_FFT_STREAM=1) against
    the oracle on full-tile shapes, both directions."""
    import jaxfun_oracle as O
    import jaxfun_b200 as jf
    monkeypatch.setenv("JFX_FFT_STREAM", "1")
    n = shape[axis]
    Vo, Vp = getattr(O, kind)(n), getattr(jf, kind)(n)
    rng = np.random.default_rng(sum(shape))
    c = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if kind == "Fourier" else 0)
    u_ref = Vo.backward(c, axis=axis)
    u = Vp.backward(torch.from_numpy(c).to(cuda), axis=axis)
    assert np.abs(u.cpu().numpy() - u_ref).max() < 1e-12 * np.abs(u_ref).max()
    f = Vp.forward(torch.from_numpy(u_ref).to(cuda), axis=axis)
    assert np.abs(f.cpu().numpy() - c).max() < 1e-11 * np.abs(c).max()
