#!/usr/bin/env python
"""Benchmark of the tensor-product transform hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--no-e2e] [--no-cpu] [--no-legs]

N = 1 (default): BASELINE configs[1] — TensorProduct forward + backward of Legendre^3 and Chebyshev^3 at 256^3 fp64.
One step = 4 transforms (Legendre backward, forward; Chebyshev backward, forward).  The same line carries, as `legs`, the
other single-GPU configs of BASELINE.json measured the same way (CUDA events, inputs resident and larger than L2):
C3 (65 536 lines of N = 1024: Fourier, Legendre, KdV nonlinear term), C4 (Cahn-Hilliard nonlinear term and one ETDRK4 step
on Fourier^2 4096^2) and C5 on ONE GPU (Legendre^3 512^3, the strong-scaling denominator of the N > 1 runs).
N > 1 (under torchrun): BASELINE configs[4] — Legendre^3 512^3 slab-decomposed over N ranks with the all-to-all exchange;
one step = backward + forward of the global field (strong scaling).  The result is CHECKED in the run: round trip against
the input and rank 0's block against the same transform on one GPU; `parallel_efficiency` = T(1 GPU, same problem, same
run) / (N * T(N)).

Prints ONE JSON line (rank 0).
  value         whole-job transforms/s with inputs resident in HBM (device-timed, max over ranks)
  e2e           the same through the public API with pinned HOST buffers: every step copies its inputs host -> device and
                its results device -> host inside the timed region (two steps in flight; `e2e_sync` = host sync per step)
  roofline      the dominant kernel (FP64 tensor-core contraction): SURVEY 8(d) algorithmic flops per launch / the
                launch time, against the FP64 tensor peak calibrated live; `issued` = the multiply-adds really issued
  roofline_hbm  Chebyshev^3: 8(d) compulsory bytes (input read once + output written once per transform) / time,
                against MEASURED_PEAKS.json; `per_launch` = one axis pass
  cpu_baseline  the NumPy/SciPy oracle (the reference's algorithm restated) on the host cores
`--impl reference` times that oracle on the same workload: jax is not installable here, so the reference arm
is the port; it imports nothing of the product.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3D tensor-product fwd/bwd transforms/s (fp64)"
UNIT = "transforms/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override the cube edge (default 256 / 512)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the C3 / C4 / C5-on-one-GPU legs of the N = 1 line")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.proc, self.path = index, None, f"/tmp/jfx_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [s.strip() for s in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel):
    """dram bytes per launch (read + write) of `kernel` from the committed `ncu --set full` capture."""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            return t[kernel]["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------------
# CPU side (oracle)
# --------------------------------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs must use every host core (BLAS via threadpoolctl,
    scipy.fft via its `workers` argument in the oracle)."""
    cores = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    return cores


def oracle_spaces(n, dims=3):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import jaxfun_oracle as O
    use_all_host_threads()
    O.Jacobi.fast_backward = True  # Vandermonde matmul instead of the scan: the faster CPU form
    return (O, O.TensorProductSpace(*[O.Legendre(n) for _ in range(dims)]),
            O.TensorProductSpace(*[O.Chebyshev(n) for _ in range(dims)]))


def cpu_step(TL, TC, cL, cC):
    """One step of the workload on the host: returns #transforms."""
    n = 0
    if TL is not None:
        u = TL.backward(cL); TL.forward(u); n += 2
    if TC is not None:
        u = TC.backward(cC); TC.forward(u); n += 2
    return n


def run_reference(args):
    """--impl reference: the oracle (NumPy/SciPy restatement of the reference algorithm) on host cores."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    multi = args.gpus > 1
    n = args.size or (512 if multi else 256)
    cores = os.cpu_count() or 1
    O, TL, TC = oracle_spaces(n)
    if multi:
        TC = None
    rng = np.random.default_rng(5 if multi else 2)
    cL = rng.standard_normal((n, n, n))
    cC = rng.standard_normal((n, n, n)) if TC is not None else None
    for _ in range(min(args.warmup, 1)):
        cpu_step(TL, TC, cL, cC)
    t0 = time.perf_counter()
    ntr = 0
    for _ in range(args.steps):
        ntr += cpu_step(TL, TC, cL, cC)
    dt = time.perf_counter() - t0
    val = ntr / dt
    sample = f"full workload: {ntr // args.steps} transforms of {n}^3 fp64 per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(n, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "NumPy/SciPy oracle (matmul via the recurrence / scipy.fft), not XLA: jax is not installable here"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["stock_scan_backward"] = stock_scan_sample(O, min(n, 256))
    except Exception as e:  # pragma: no cover
        line["stock_scan_backward"] = {"error": f"{type(e).__name__}: {e}"}
    emit(line)


def _slab_env():
    """The exchange switches of jaxfun_b200/sharding.py, read from the environment (the reference arm must not import the product)."""
    def flag(name, default="0"):
        return os.environ.get(name, default) != "0"
    try:
        chunks = max(1, int(os.environ.get("JFX_SLAB_CHUNKS", "1")))
    except ValueError:
        chunks = 1
    return {"slab_chunks": chunks, "slab_p2p": flag("JFX_SLAB_P2P", "1"), "slab_fused_pack": flag("JFX_SLAB_FUSED_PACK"),
            "slab_native": flag("JFX_SLAB_NATIVE", "1")}


def workload_config(n, gpus):
    if gpus > 1:
        cfg = _slab_env()
        cfg.update({"workload": f"C5: Legendre^3 {n}^3 fp64 slab-decomposed (axis0<->axis1 all-to-all), backward+forward per step",
                    "shape": [n, n, n], "parallelism": f"slab{gpus}", "l2": "arrays (>= 134 MB per rank) larger than L2",
                    "scaling_note": "strong scaling of ONE 512^3 problem; its one-GPU time is measured in the same run "
                                    "(single_gpu_same_size) and also carried by the N=1 line (legs.c5_single_gpu); the N=1 "
                                    "headline itself is BASELINE configs[1] (256^3 mix), a different workload"})
        return cfg
    return {"workload": f"C2: TensorProduct backward+forward, Legendre^3 and Chebyshev^3, {n}^3 fp64 (4 transforms/step)",
            "shape": [n, n, n], "parallelism": "single", "l2": f"inputs larger than L2 ({8 * n**3 / 1e6:.0f} MB arrays, 4 buffers per transform)",
            "scaling_note": "N>1 runs BASELINE configs[4] (Legendre^3 512^3, strong scaling); its one-GPU denominator is legs.c5_single_gpu"}


def stock_scan_sample(O, n, lines=2048):
    """The reference's OWN backward for Legendre/Jacobi is an N-step recurrence scan (Jacobi.py:65-110), not a matmul.  Timed
    here on a bounded sample (`lines` lines of one axis pass) and reported beside the matmul form the arm uses."""
    import numpy as np
    V = O.Legendre(n)
    c = np.random.default_rng(7).standard_normal((lines, n))
    prev = O.Jacobi.fast_backward
    try:
        O.Jacobi.fast_backward = False
        t0 = time.perf_counter(); V.backward(c, axis=-1); t_scan = time.perf_counter() - t0
        O.Jacobi.fast_backward = True
        V.backward(c, axis=-1)
        t0 = time.perf_counter(); V.backward(c, axis=-1); t_mm = time.perf_counter() - t0
    finally:
        O.Jacobi.fast_backward = prev
    per_transform = 3 * n * n / lines              # axis passes x lines per pass of one n^3 backward
    return {"sample": f"{lines} lines of N = {n} (1/{n * n // lines} of one axis pass of a {n}^3 backward), one run each",
            "scan_lines_per_s": lines / t_scan, "matmul_lines_per_s": lines / t_mm, "scan_over_matmul_time": t_scan / t_mm,
            "scan_s_per_backward_extrapolated": t_scan * per_transform,
            "note": "stock algorithm of the reference (recurrence scan, vector ops per step) vs the Vandermonde matmul this arm "
                    "times; the extrapolation is lines x sample time, not a measurement of a full transform"}


# --------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------
def run_legs(jf, L, torch, dev, fp64_peak, hbm):
    """The other single-GPU configs of BASELINE.json, measured like the headline (CUDA events, warm, inputs resident and larger
    than L2): C3 batched 1-D (65 536 x 1024), C4 Cahn-Hilliard 4096^2, C5 on one GPU.  Each entry carries its own roofline
    fraction by the SURVEY 8(d) count."""
    import numpy as np
    import sympy as sp
    from jaxfun_b200.integrators import ETDRK4, NonlinearTerm, field

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def hbm_entry(ms, nbytes, launches=None, fft_flops=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        d = {"ms": ms, "compulsory_bytes": nbytes, "achieved_gbs": gbs, "frac_of_hbm": gbs / hbm}
        if launches is not None:
            d["launches"] = launches
        if fft_flops is not None:
            # second bound of the fused nonlinear terms: the butterflies themselves (5 n log2 n real operations per 1-D FFT
            # line).  They are additions and multiplications, hardly any fused multiply-adds, so the FP64 pipe retires at
            # most HALF of its FMA peak on them; the floor printed here is flops / (peak / 2) beside bytes / HBM peak.
            tf = fft_flops / (ms * 1e-3) / 1e12
            d["fft"] = {"flops": fft_flops, "achieved_tflops": tf, "frac_of_fp64_peak": tf / fp64_peak,
                        "floor_ms_fp64_non_fma": fft_flops / (fp64_peak * 0.5e12) * 1e3,
                        "floor_ms_hbm": nbytes / (hbm * 1e9) * 1e3,
                        "binding": "fp64" if fft_flops / (fp64_peak * 0.5e12) > nbytes / (hbm * 1e9) else "hbm"}
        return d

    legs = {}
    g = torch.Generator(device=dev).manual_seed(3)
    # ---- C3: 65 536 independent lines of N = 1024 ------------------------------------------------------------------------
    rows, n3 = 65536, 1024
    try:
        F = jf.Fourier(n3, domain=(-30.0, 30.0))
        cF = torch.view_as_complex(torch.randn(rows, n3, 2, dtype=torch.float64, device=dev, generator=g)) * 0.05
        uF = F.backward(cF)
        bF = 2.0 * 16 * rows * n3
        c3 = {"shape": [rows, n3],
              "fourier_backward": hbm_entry(timed(lambda: F.backward(cF)), bF, 1),
              "fourier_forward": hbm_entry(timed(lambda: F.forward(uF)), bF, 1)}
        del uF
        u, (x,) = field(F)
        term = NonlinearTerm(F, -u * u.diff(x))
        term(cF)
        c3["kdv_nonlinear"] = hbm_entry(timed(lambda: term(cF)), bF, term.launches(cF),
                                        fft_flops=3 * rows * 5.0 * n3 * np.log2(n3))      # 2 inverse + 1 forward FFT per line
        c3["kdv_nonlinear"]["note"] = "forward(-(u u_x)): one coefficient read + one coefficient write are compulsory (8d)"
        del term, cF
        torch.cuda.empty_cache()
        Lg = jf.Legendre(n3)
        cL = torch.randn(rows, n3, dtype=torch.float64, device=dev, generator=g)
        uL = Lg.backward(cL)
        fl = 2.0 * rows * n3 * n3
        plans = {k[0]: p for k, p in Lg._plans.items()}
        for name, fn, op in (("legendre_backward", lambda: Lg.backward(cL), L.OP_BACKWARD),
                             ("legendre_forward", lambda: Lg.forward(uL), L.OP_FORWARD)):
            if op == L.OP_FORWARD:
                fn()
                plans = {k[0]: p for k, p in Lg._plans.items()}
            ms = timed(fn)
            issued = plans[op].flops_executed if op in plans else fl
            c3[name] = {"ms": ms, "algorithmic_flops": fl, "achieved_tflops": fl / (ms * 1e-3) / 1e12,
                        "frac_of_fp64_tensor_peak": fl / (ms * 1e-3) / 1e12 / fp64_peak,
                        "issued_tflops": issued / (ms * 1e-3) / 1e12, "issued_frac": issued / (ms * 1e-3) / 1e12 / fp64_peak}
        del cL, uL, Lg
        legs["c3_batched_1d"] = c3
    except Exception as e:  # pragma: no cover
        legs["c3_batched_1d"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    # ---- C4: Cahn-Hilliard on Fourier^2 4096^2 ---------------------------------------------------------------------------
    try:
        n4, dom = 4096, (0.0, 1.0)
        T4 = jf.TensorProduct(jf.Fourier(n4, domain=dom), jf.Fourier(n4, domain=dom))
        uu, (x, y) = field(T4)
        term4 = NonlinearTerm(T4, -(6 * uu * (uu.diff(x) ** 2 + uu.diff(y) ** 2) + 3 * uu**2 * (uu.diff(x, 2) + uu.diff(y, 2))))
        uh = 1e-2 * torch.view_as_complex(torch.randn(n4, n4, 2, dtype=torch.float64, device=dev, generator=g))
        kx = torch.from_numpy(np.asarray(T4.basespaces[0].wavenumbers(), dtype=float) * 2 * np.pi).to(dev)
        k2 = kx[:, None] ** 2 + kx[None, :] ** 2
        uh = uh / (1.0 + k2 / (2 * np.pi) ** 2) ** 1.5            # smooth field (as tests/test_at_size_gpu.py)
        term4(uh)
        b4 = 2.0 * 16 * n4 * n4
        c4 = {"shape": [n4, n4], "nonlinear_N": hbm_entry(timed(lambda: term4(uh)), b4, term4.launches(uh),
                                                        # 5 inverse + 1 forward 2-D FFT = 12 x 4096 lines of n = 4096
                                                        fft_flops=6 * 2 * n4 * 5.0 * n4 * np.log2(n4))}
        c4["nonlinear_N"]["note"] = "-(6u(ux^2+uy^2)+3u^2(uxx+uyy)), 5 leaves: one coefficient read + one write are compulsory (8d)"
        Ld = (-(k2 * k2) * 1e-4 - 0 * k2).to(torch.complex128)   # -gamma k^4 - alpha... diagonal linear operator of the example
        integ = ETDRK4(T4, linear_diag=Ld, nonlinear=term4)
        dt = 5e-2 / 320
        integ.setup(dt)
        ms_step = timed(lambda: integ.step(uh, dt), reps=3)
        c4["etdrk4_step"] = {"ms": ms_step, "nonlinear_evaluations": 4, "diagonal_combinations": 5,
                             "compulsory_bytes": 4 * b4 + 5 * 4 * 16.0 * n4 * n4,
                             "achieved_gbs": (4 * b4 + 5 * 4 * 16.0 * n4 * n4) / (ms_step * 1e-3) / 1e9,
                             "frac_of_hbm": (4 * b4 + 5 * 4 * 16.0 * n4 * n4) / (ms_step * 1e-3) / 1e9 / hbm,
                             "note": "4 nonlinear terms (coefficient in + out each) + 5 fused diagonal combinations (about 3 "
                                     "fields in + 1 out each); dt = 5e-2 / 320 as in examples/cahn_hilliard2D_etdrk4.py"}
        legs["c4_cahn_hilliard"] = c4
        del term4, integ, uh, T4, Ld, k2
    except Exception as e:  # pragma: no cover
        legs["c4_cahn_hilliard"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    # ---- mixed product (complex data on a polynomial axis): Fourier x Fourier x Legendre 256^3 c128 ------------
    try:
        nm = 256
        Tm = jf.TensorProduct(jf.Fourier(nm), jf.Fourier(nm), jf.Legendre(nm))
        cm = torch.view_as_complex(torch.randn(nm, nm, nm, 2, dtype=torch.float64, device=dev, generator=g))
        um = Tm.backward(cm)
        msb, msf = timed(lambda: Tm.backward(cm)), timed(lambda: Tm.forward(um))
        legs["mixed_FxFxLeg_256"] = {"shape": [nm] * 3, "dtype": "c128", "backward_ms": msb, "forward_ms": msf,
                                     "launches": [p.launches for p in Tm._plans.values()],
                                     "note": "two FFT passes (HBM-bound, 537 MB each) + one real-table contraction of complex lines "
                                             "(CPLX_NT: 2 x 8.59 GFLOP algorithmic, unfolded) per transform"}
        del Tm, cm, um
    except Exception as e:  # pragma: no cover
        legs["mixed_FxFxLeg_256"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    # ---- Fourier x Fourier x polynomial Helmholtz solve: 65 536 pentadiagonal-type systems of n = 254 ---
    try:
        from jaxfun_b200.galerkin.tpsolve import WavenumberBandedSolver
        nb, nsys = 254, 256 * 256
        Pb = np.zeros((2, 3, nb))
        Pb[0, 1] = 4.0 + 0.01 * np.arange(nb)
        Pb[1, 0, :nb - 2], Pb[1, 1], Pb[1, 2, 2:] = -0.2, 1.0, -0.2
        Wb = np.stack([np.ones(nsys), 1.0 + np.random.default_rng(6).random(nsys)])
        out_b = {}
        for tag, shape, pa in (("poly_axis_last", (256, 256, nb), 2), ("poly_axis_middle", (256, nb, 256), 1)):
            Sb = WavenumberBandedSolver(pa, shape, Wb, Pb, (-2, 0, 2))
            rb = [torch.view_as_complex(torch.randn(*shape, 2, dtype=torch.float64, device=dev, generator=g)) for _ in range(2)]
            ob = torch.empty_like(rb[0])
            cnt = [0]

            def solve_once():
                cnt[0] += 1
                Sb.solve(rb[cnt[0] % 2], out=ob)          # two 266 MB inputs alternate: nothing of the input is L2-resident
            ms = timed(solve_once, reps=6)
            nbytes = 2 * rb[0].numel() * 16 + 5 * 8 * nb * nsys
            out_b[tag] = dict(hbm_entry(ms, nbytes, launches=1), shape=list(shape))
            del Sb, rb, ob
        out_b["note"] = ("jfx_banded_solve, complex128 right-hand sides, real factors with offsets (-2, 0, 2): compulsory bytes = "
                         "right-hand side in + solution out + factored band read once (the y round trip of L y = b is extra traffic)")
        legs["banded_FxFxPoly_256"] = out_b
    except Exception as e:  # pragma: no cover
        legs["banded_FxFxPoly_256"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    # ---- C5 on ONE GPU: the strong-scaling denominator of the N > 1 runs -------------------------------------------------
    try:
        n5 = 512
        T5 = jf.TensorProduct(*[jf.Legendre(n5) for _ in range(3)])
        c5 = torch.randn(n5, n5, n5, dtype=torch.float64, device=dev, generator=g)
        ms5 = timed(lambda: T5.forward(T5.backward(c5)), reps=3)
        fl5 = 2 * 6.0 * float(n5) ** 4
        pl = [p for p in T5._plans.values()]
        issued = sum(p.flops_executed for p in pl) if len(pl) == 2 else fl5
        legs["c5_single_gpu"] = {"shape": [n5] * 3, "ms_per_step": ms5, "value": 2.0 / (ms5 * 1e-3), "unit": UNIT,
                                 "achieved_tflops": fl5 / (ms5 * 1e-3) / 1e12,
                                 "frac_of_fp64_tensor_peak": fl5 / (ms5 * 1e-3) / 1e12 / fp64_peak,
                                 "issued_frac": issued / (ms5 * 1e-3) / 1e12 / fp64_peak,
                                 "note": "Legendre^3 512^3 backward + forward on one GPU: T(1) of parallel_efficiency = T(1) / (N T(N))"}
        del T5, c5
    except Exception as e:  # pragma: no cover
        legs["c5_single_gpu"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return legs


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if multi:
        # keep stdout for the ONE JSON line: NCCL's own messages (version banner, NCCL_DEBUG output) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n = args.size or (512 if multi else 256)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    lib = L.load()
    import ctypes as C

    if not multi:
        TL = jf.TensorProduct(*[jf.Legendre(n) for _ in range(3)])
        TC = jf.TensorProduct(*[jf.Chebyshev(n) for _ in range(3)])
        g = torch.Generator(device=dev).manual_seed(2)
        cL = torch.randn(n, n, n, dtype=torch.float64, device=dev, generator=g)
        cC = torch.randn(n, n, n, dtype=torch.float64, device=dev, generator=g)
        pLb, pLf = TL._plan(L.OP_BACKWARD, cL), TL._plan(L.OP_FORWARD, cL)
        pCb, pCf = TC._plan(L.OP_BACKWARD, cC), TC._plan(L.OP_FORWARD, cC)
        uL, uC = torch.empty_like(cL), torch.empty_like(cC)
        oL, oC = torch.empty_like(cL), torch.empty_like(cC)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

        def step(i=None):
            if i is not None:
                evs[i][0].record()
            pLb.execute(cL, uL); pLf.execute(uL, oL)
            if i is not None:
                evs[i][1].record()
            pCb.execute(cC, uC); pCf.execute(uC, oC)
            if i is not None:
                evs[i][2].record()
            return 4

        launches_per_step = pLb.launches + pLf.launches + pCb.launches + pCf.launches
        flops_L = pLb.flops + pLf.flops
        flops_L_exec = pLb.flops_executed + pLf.flops_executed
        bytes_C = pCb.bytes + pCf.bytes
    else:
        from jaxfun_b200.sharding import SlabTensorProduct
        T = jf.TensorProduct(*[jf.Legendre(n) for _ in range(3)])
        S = SlabTensorProduct(T)
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        c_loc = torch.randn(n // world, n, n, dtype=torch.float64, device=dev, generator=g)

        def step(i=None):
            u = S.backward(c_loc)
            S.forward(u)
            return 2

        # ---- the result is checked before anything is timed: round trip, and rank 0's physical block against the SAME
        # global transform on one GPU (every rank can regenerate every block: seeds are 5 + rank)
        u_loc = S.backward(c_loc)
        c_back = S.forward(u_loc)
        rt = float((c_back - c_loc).abs().max() / c_loc.abs().max())
        check = {"roundtrip_rel_err": rt}
        if rank == 0:
            blocks = [torch.randn(n // world, n, n, dtype=torch.float64, device=dev,
                                  generator=torch.Generator(device=dev).manual_seed(5 + r)) for r in range(world)]
            c_glob = torch.cat(blocks, dim=0)
            del blocks
            u_glob = T.backward(c_glob)
            b = n // world
            ref_blk = u_glob[:, :b, :]                      # physical sharding: axis 1, rank 0 owns the first n/P columns
            check["rank0_block_vs_single_gpu_rel_err"] = float((u_loc - ref_blk).abs().max() / u_glob.abs().max())
            del c_glob, u_glob, ref_blk
        tt = torch.tensor([rt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        check["roundtrip_rel_err_max_over_ranks"] = float(tt.item())
        ok = check["roundtrip_rel_err_max_over_ranks"] < 1e-11 and check.get("rank0_block_vs_single_gpu_rel_err", 0.0) < 1e-12
        check["ok"] = bool(ok)
        del u_loc, c_back
        torch.cuda.empty_cache()

        launches_per_step = 2 * 3 + 2  # 3 contraction passes + 1 repack per transform (+ NCCL)
        flops_L = 2 * 6.0 * float(n) ** 4 / world
        flops_L_exec = flops_L
        bytes_C = 0.0

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    ntr = 0
    for i in range(args.steps):
        ntr += step(i if not multi else None)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if multi:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # nvidia-smi needs ~1 s to deliver samples: if the timed region was shorter, keep running the SAME
    # steps (untimed) so the clock/throttle record reflects this workload under load.  The number of extra
    # steps is derived from the all-reduced time, so every rank issues the same collectives.
    n_extra = 0
    if ms < 2000.0:
        n_extra = min(2000, int((2000.0 - ms) / max(ms / args.steps, 1e-3)) + 1)
    for _ in range(n_extra):
        step()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    value = ntr / (ms * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, world), "gpu_launches": launches_per_step * args.steps,
    }

    peaks = measured_peaks()
    # FP64 denominators: calibrated live (no FP64 figure in MEASURED_PEAKS.json)
    dm, df = C.c_double(), C.c_double()
    L.check(lib.jfx_calibrate_dmma(None, 4000, C.byref(dm)))
    L.check(lib.jfx_calibrate_dfma(None, 4000, C.byref(df)))
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); torch.matmul(a, b); c1.record(); torch.cuda.synchronize()
    cublas_tf = 2 * 8192.0**3 / (c0.elapsed_time(c1) * 1e-3) / 1e12
    del a, b
    fp64_peak = max(dm.value, df.value, cublas_tf)

    if not multi:
        tL = sum(evs[i][0].elapsed_time(evs[i][1]) for i in range(args.steps)) / args.steps
        tC = sum(evs[i][1].elapsed_time(evs[i][2]) for i in range(args.steps)) / args.steps
        n_l = pLb.launches + pLf.launches
        ach = flops_L / (tL * 1e-3) / 1e12
        ach_exec = flops_L_exec / (tL * 1e-3) / 1e12
        folded = flops_L_exec < 0.75 * flops_L
        kname = ("dgemm_dmma_fold (parity-folded FP64 tensor-core per-axis Vandermonde contraction: half the multiply-adds, "
                 "TMA-fed mbarrier pipeline)") if folded else \
                "dgemm_dmma_tma (FP64 tensor-core per-axis Vandermonde contraction, TMA-fed mbarrier pipeline)"
        fl_launch_alg, fl_launch_iss = flops_L / n_l, flops_L_exec / n_l
        line["roofline"] = {
            "kernel": kname, "bound": "tensor",
            # ALGORITHMIC flops per launch (2 N Nq per line) / average launch time.  The folded kernel issues
            # only half of them (mirror symmetry of the table), so this figure can exceed the tensor peak: `issued` is the
            # tensor-pipe utilisation (<= 1 by construction) and the number that measures kernel quality.
            "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
            "issued": {"achieved": ach_exec, "frac": ach_exec / fp64_peak, "flops_per_launch": fl_launch_iss},
            "algorithmic_flops_per_launch": fl_launch_alg,
            "note": ("frac > 1 is possible and legitimate here: the 8(d) count assumes a full dense contraction, the folded kernel "
                     "needs half the multiply-adds for mirror-symmetric tables (plus correction k-tiles for the highest modes "
                     "from n = 320 on); kernel quality = issued.frac") if folded else "issued = algorithmic flops (no folding)",
            "traffic": ncu_traffic("dgemm_dmma_fold" if folded else "dgemm_dmma"), "launches_per_step": n_l,
            "avg_launch_ms": tL / n_l,
            "peak_source": "live calibration: max(register-resident DMMA loop [best of 1/2/8 CTAs per SM], DFMA loop, "
                           "cuBLAS DGEMM 8192^3); MEASURED_PEAKS.json has no FP64 figure",
            "fp64_calibration_tflops": {"dmma_regs": dm.value, "dfma_regs": df.value, "cublas_dgemm_8192": cublas_tf,
                                        "nominal": 37.0},
        }
        hbm = peaks.get("hbm_gbs", 6650.0)
        n_c = pCb.launches + pCf.launches
        achC = bytes_C / (tC * 1e-3) / 1e9                 # whole transforms: compulsory bytes / time
        per_launch_bytes = 2.0 * 8 * n**3                   # one axis pass reads the field once and writes it once
        ach_launch = per_launch_bytes / (tC / n_c * 1e-3) / 1e9
        # context for the HBM fraction: what a plain device copy of ONE field of this size reaches (the
        # MEASURED_PEAKS figure is a 2 GiB copy; a 134 MB launch also pays its ramp-up and tail)
        cp_src, cp_dst = torch.empty_like(cC), torch.empty_like(cC)
        for _ in range(3):
            cp_dst.copy_(cp_src)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(10):
            cp_dst.copy_(cp_src)
        k1.record()
        torch.cuda.synchronize()
        copy_gbs = per_launch_bytes / (k0.elapsed_time(k1) / 10 * 1e-3) / 1e9
        del cp_src, cp_dst
        line["roofline_hbm"] = {
            "kernel": "fft2_kernel (Chebyshev DCT axis pass; Chebyshev^3 backward+forward = 6 launches)",
            "bound": "hbm",
            # compulsory bytes (input read once + output written once per 3-D transform) / time
            "achieved": achC, "peak": hbm, "unit": "GB/s", "frac": achC / hbm,
            "algorithmic_bytes_per_step": bytes_C,
            "per_launch": {"achieved": ach_launch, "frac": ach_launch / hbm, "bytes_per_launch": per_launch_bytes,
                           "avg_launch_ms": tC / n_c, "copy_same_size_gbs": copy_gbs,
                           "frac_of_copy_same_size": ach_launch / copy_gbs,
                           "note": "one axis pass: the field read once and written once"},
            "traffic": ncu_traffic("fft2_kernel"), "launches_per_step": n_c,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650",
            "note": "three axis passes over a 134 MB field (> L2) cap the transform-level fraction at 1/3 of the per-launch one",
        }
        line["detail"] = {
            "legendre3": {"ms_per_pair": tL, "transforms_per_s": 2e3 / tL, "tflops": ach},
            "chebyshev3": {"ms_per_pair": tC, "transforms_per_s": 2e3 / tC, "gbs_compulsory": achC},
        }
        # live A/B of the parity folding on the Legendre pair (outside the timed region; context for `roofline`):
        # the same plans created with JFX_DMMA_FOLD=0 run the plain tensor-core kernel dgemm_dmma_tma
        try:
            prev = os.environ.get("JFX_DMMA_FOLD")
            os.environ["JFX_DMMA_FOLD"] = "0"
            try:
                TLp = jf.TensorProduct(*[jf.Legendre(n) for _ in range(3)])
                qb, qf = TLp._plan(L.OP_BACKWARD, cL), TLp._plan(L.OP_FORWARD, cL)
            finally:
                if prev is None:
                    os.environ.pop("JFX_DMMA_FOLD", None)
                else:
                    os.environ["JFX_DMMA_FOLD"] = prev
            for _ in range(3):
                qb.execute(cL, uL); qf.execute(uL, oL)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(10):
                qb.execute(cL, uL); qf.execute(uL, oL)
            a1.record()
            torch.cuda.synchronize()
            t_plain = a0.elapsed_time(a1) / 10
            ref_o = oL.clone()
            pLb.execute(cL, uL); pLf.execute(uL, oL)
            torch.cuda.synchronize()
            line["fold_ab"] = {
                "legendre3_ms_per_pair_plain": t_plain, "legendre3_ms_per_pair_folded": tL, "speedup": t_plain / tL,
                "plain_tflops": (qb.flops + qf.flops) / (t_plain * 1e-3) / 1e12,
                "plain_frac_of_peak": (qb.flops + qf.flops) / (t_plain * 1e-3) / 1e12 / fp64_peak,
                "max_rel_diff_folded_vs_plain": float((oL - ref_o).abs().max() / ref_o.abs().max()),
                "note": "same Legendre^3 backward+forward with JFX_DMMA_FOLD=0 at plan creation (dgemm_dmma_tma), 10 iterations, CUDA events",
            }
            del qb, qf, TLp, ref_o
        except Exception as e:  # pragma: no cover  (context only: never fail the bench line over it)
            line["fold_ab"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_legs and not args.size:
            line["legs"] = run_legs(jf, L, torch, dev, fp64_peak, hbm)
    else:
        ach = 2 * 6.0 * float(n) ** 4 / world / (ms / args.steps * 1e-3) / 1e12
        # share of the algorithmic flops the local plans actually issue (0.5 when every pass is parity-folded)
        issued_ratio = 1.0
        try:
            from jaxfun_b200.engine import Plan as _Plan
            plans = [p for be in S._backends.values() for p in be._plans.values() if isinstance(p, _Plan)]
            fa, fe = sum(p.flops for p in plans), sum(p.flops_executed for p in plans)
            if fa > 0:
                issued_ratio = fe / fa
        except Exception:
            pass
        line["roofline"] = {"kernel": "dgemm_dmma_fold / dgemm_dmma_tma inside the slab transform (per rank, incl. exchange time)",
                            "bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                            "issued": {"achieved": ach * issued_ratio, "frac": ach * issued_ratio / fp64_peak},
                            "traffic": None,
                            "note": "achieved = SURVEY 8(d) algorithmic flops per rank (6 N^4 per transform / ranks) over the whole "
                                    "step time including the exchange; issued = what the parity-folded passes really issue",
                            "peak_source": "live calibration (see N=1 line)"}
        line["check"] = check
        # NVLink side of the exchange: each rank sends (P-1)/P of its block in each of the two transforms of a step
        sent = 2.0 * 8.0 * n**3 / world * (world - 1) / world
        p2p_used = any(isinstance(v, dict) and "hdls" in v for be in S._backends.values() for v in be._plans.values())
        p2p_err = next((be.p2p_error for be in S._backends.values() if getattr(be, "p2p_error", None)), None)
        native = [v for be in S._backends.values() for k, v in be._plans.items()
                  if isinstance(k, tuple) and k and k[0] == "native" and isinstance(v, dict)]
        if native:
            path = ("jfx_slab_execute (one C-ABI call per transform): " +
                    ("peer stores from the epilogue of the last local pass" if all(v["fused"] for v in native)
                     else "strided peer copies") + " + device-side flag barrier over symmetric memory (NVLink), no NCCL collective "
                    "and no torch op on the data path")
        elif p2p_used:
            path = ("peer stores from the epilogue of the last local pass (symmetric memory over NVLink), barrier by "
                    "torch symmetric memory, no NCCL collective on the data path")
        else:
            path = "jfx_slab_pack + NCCL all_to_all_single (+ unpack)"
        line["exchange"] = {"bytes_sent_per_rank_per_step": sent, "mode": _slab_env(), "path": path,
                            "p2p_fallback_reason": p2p_err}
        if rank == 0:
            # same global problem on ONE GPU: the strong-scaling denominator, measured in this run
            try:
                T1 = jf.TensorProduct(*[jf.Legendre(n) for _ in range(3)])
                c1g = torch.randn(n, n, n, dtype=torch.float64, device=dev)
                for _ in range(2):
                    T1.forward(T1.backward(c1g))
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); s0.record()
                for _ in range(3):
                    T1.forward(T1.backward(c1g))
                s1.record(); torch.cuda.synchronize()
                ms1 = s0.elapsed_time(s1) / 3
                line["single_gpu_same_size"] = {
                    "ms_per_step": ms1, "value": 2.0 / (ms1 * 1e-3), "unit": UNIT,
                    "note": "the SAME 512^3 backward+forward on one GPU, measured on rank 0 in this run: the strong-"
                            "scaling denominator (the default N=1 bench line is a different workload: 256^3 mix)"}
                line["parallel_efficiency"] = ms1 / (world * (ms / args.steps))
                line["exchange"]["local_compute_ms_per_step_ideal"] = ms1 / world
                line["exchange"]["non_overlapped_ms_per_step"] = ms / args.steps - ms1 / world
            except Exception as e:  # pragma: no cover
                line["single_gpu_same_size"] = {"error": str(e)}
        if not check["ok"]:
            line["INVALID"] = "the slab transform failed its result check (see `check`)"

    if rank == 0:
        line["clocks"] = clocks

    # ---- e2e: public API, pinned host buffers, H2D + D2H inside the timed region ---------------
    if not args.no_e2e and not multi:
        # One step = the same 4 transforms.  Every step copies the step's INPUTS (the two coefficient
        # arrays) from pinned host memory to the device, calls the public TensorProductSpace API
        # (backward, then forward on its result — the call sequence of the reference's round-trip tests)
        # and copies the step's RESULTS (the two coefficient arrays that come back) to pinned host memory.
        rng = np.random.default_rng(2)
        h_in = [torch.from_numpy(rng.standard_normal((n, n, n))).pin_memory() for _ in range(2)]
        h_out = [torch.empty(n, n, n, dtype=torch.float64).pin_memory() for _ in range(2)]
        k_e2e = max(3, min(args.steps, 10))

        # The two bases of a step are independent, so each gets its own CUDA stream: the H2D copy of one
        # overlaps the transforms of the other and (PCIe being full duplex) the D2H copy of its results.
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

        def e2e_step():
            for T, hi, ho, st in ((TL, h_in[0], h_out[0], streams[0]), (TC, h_in[1], h_out[1], streams[1])):
                with torch.cuda.stream(st):
                    c = hi.to(dev, non_blocking=True)
                    o = T.forward(T.backward(c))
                    ho.copy_(o, non_blocking=True)
            for st in streams:
                st.synchronize()                 # every step ends with its results in host memory
            return 4
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ntr = 0
        for _ in range(k_e2e):
            ntr += e2e_step()
        dt = time.perf_counter() - t0
        e2e_sync = {"value": ntr / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n**3,
                    "d2h_bytes_per_step": 2 * 8 * n**3, "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e,
                    "api": "as e2e, with a host synchronisation at the end of EVERY step (no two steps in flight)"}
        e2e_sync["roundtrip_max_abs_err"] = float(max((h_out[i] - h_in[i]).abs().max().item() for i in range(2)))
        # The e2e figure: the same loop with two steps in flight (double-buffered pinned outputs).  The host only waits
        # for step i - 1 before it submits step i + 1, and for everything at the end, so the H2D copies of one step
        # overlap the D2H copies of the previous one (PCIe is full duplex).  Every step still copies its inputs in and its
        # results out inside the timed region.
        h_out2 = [torch.empty(n, n, n, dtype=torch.float64).pin_memory() for _ in range(2)]
        outs = [h_out, h_out2]
        done = [None, None]
        # three stages on three streams: host -> device copies, the four transforms, device -> host copies.  A stream per
        # basis (as in e2e_sync) serialises the H2D copy of step i + 1 behind the D2H copy of step i; with a copy stream in
        # each direction both PCIe directions stay busy and the step costs max(H2D, D2H, compute) instead of their sum.
        s_in, s_cmp, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        d_in = [[torch.empty(n, n, n, dtype=torch.float64, device=dev) for _ in range(2)] for _ in range(2)]
        ev_cmp = [None, None]                      # compute of the step that last used device slot k has finished

        def submit(i):
            k = i & 1
            with torch.cuda.stream(s_in):
                if ev_cmp[k] is not None:
                    s_in.wait_event(ev_cmp[k])     # the transforms of step i - 2 have read d_in[k]
                for b in range(2):
                    d_in[k][b].copy_(h_in[b], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(e_in)
                res = [T.forward(T.backward(d_in[k][b])) for b, T in enumerate((TL, TC))]     # the public API
                e_c = torch.cuda.Event()
                e_c.record(s_cmp)
                ev_cmp[k] = e_c
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_c)
                for b in range(2):
                    res[b].record_stream(s_out)
                    outs[k][b].copy_(res[b], non_blocking=True)
                e_o = torch.cuda.Event()
                e_o.record(s_out)
            return [e_o]
        k_pipe = 2 * k_e2e
        for i in range(2):
            done[i & 1] = submit(i)
        torch.cuda.synchronize()
        done = [None, None]
        t0 = time.perf_counter()
        for i in range(k_pipe):
            if done[i & 1] is not None:          # the output buffers of step i - 2 are about to be reused
                for e in done[i & 1]:
                    e.synchronize()
            done[i & 1] = submit(i)
        torch.cuda.synchronize()
        dtp = time.perf_counter() - t0
        line["e2e"] = {"value": 4 * k_pipe / dtp, "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * n**3,
                       "d2h_bytes_per_step": 2 * 8 * n**3, "steps": k_pipe, "ms_per_step": 1e3 * dtp / k_pipe,
                       "api": "TensorProductSpace.backward/forward on device tensors; per step: 2 coefficient arrays pinned host -> "
                              "device, 4 transforms, 2 result arrays device -> pinned host; three CUDA streams (H2D, transforms, "
                              "D2H), two steps in flight (double-buffered device inputs and pinned results; the host waits for the "
                              "results of step i - 2 before it reuses their buffers)",
                       "roundtrip_max_abs_err": float(max((outs[(k_pipe - 1) & 1][i] - h_in[i]).abs().max().item() for i in range(2)))}
        line["e2e_sync"] = e2e_sync
        del h_out2
        # the same through the C-ABI host-pointer entry (jfx_execute_host): EVERY transform host -> host
        hin = jf.PinnedArray((n, n, n), np.float64)
        hmid = jf.PinnedArray((n, n, n), np.float64)
        hout = jf.PinnedArray((n, n, n), np.float64)
        hin.array[...] = h_in[0].numpy()

        def host_step():
            for T in (TL, TC):
                pb, pf = T._plan(L.OP_BACKWARD, cL), T._plan(L.OP_FORWARD, cL)
                pb.execute_host(hin.array, hmid.array)
                pf.execute_host(hmid.array, hout.array)
            return 4
        host_step()
        t0 = time.perf_counter()
        ntr = 0
        for _ in range(3):
            ntr += host_step()
        dt = time.perf_counter() - t0
        line["e2e_host_calls"] = {"value": ntr / dt, "unit": UNIT, "h2d_bytes_per_step": 4 * 8 * n**3,
                                  "d2h_bytes_per_step": 4 * 8 * n**3, "steps": 3, "ms_per_step": 1e3 * dt / 3,
                                  "api": "jfx_execute_host (C ABI, host pointer in -> host pointer out) per transform"}
    elif multi:
        # e2e for the slab path: each rank stages its block from pinned host memory and reads it back
        hin = torch.empty(n // world, n, n, dtype=torch.float64).pin_memory()
        hout = torch.empty(n // world, n, n, dtype=torch.float64).pin_memory()
        hin.normal_()
        k_e2e = 3
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            d = hin.to(dev, non_blocking=True)
            o = S.forward(S.backward(d))
            hout.copy_(o, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": 2 * k_e2e / float(tt.item()), "unit": UNIT,
                       "h2d_bytes_per_step": 8 * n**3, "d2h_bytes_per_step": 8 * n**3, "steps": k_e2e}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    if rank == 0 and not multi and not args.no_cpu:
        O, OL, OC = oracle_spaces(n)
        rng = np.random.default_rng(2)
        hL = rng.standard_normal((n, n, n))
        cpu_step(OL, OC, hL, hL)  # warm-up (table construction, FFT plans)
        t0 = time.perf_counter()
        ntr, nst = 0, 0
        while time.perf_counter() - t0 < 10.0 and nst < 8:
            ntr += cpu_step(OL, OC, hL, hL)
            nst += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ntr / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{nst} full steps ({ntr} transforms of {n}^3 fp64), {dt:.1f} s",
                                "note": "NumPy/SciPy oracle of the reference algorithm (threaded BLAS + scipy.fft workers), "
                                        "not XLA: jax is not installable in this image"}
    if rank == 0:
        emit(line)
    if multi:
        dist.destroy_process_group()


def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout() -> None:
    """Libraries (NCCL's version banner, NCCL_DEBUG output, torch warnings) write to fd 1 from C code; send
    everything that is not the result line to stderr so that stdout carries exactly one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
