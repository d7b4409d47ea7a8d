"""Which k-tile's operand was wrong?  OUT_NN pass (backward along axis 0 of [n, m]) with the folded kernel vs float64 matmul."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
dev = torch.device("cuda:0")
n, m = 256, 65536
V = jf.Legendre(n)
T = torch.from_numpy(np.ascontiguousarray(V._dense_table(L.OP_BACKWARD, n, n, 0))).to(dev)     # [j, k]
sig = torch.tensor([1.0 if k % 2 == 0 else -1.0 for k in range(n)], dtype=torch.float64, device=dev)
A = 0.5 * (T + sig[None, :] * T.flip(0))                    # symmetric part: what the kernel multiplies with (rows j < n/2)
c = torch.randn(n, m, dtype=torch.float64, device=dev)
ref = T @ c
half, kt_n = n // 2, (n // 2) // 16
found = 0
for rep in range(12):
    u = V.backward(c, axis=0); torch.cuda.synchronize()
    d = u - ref
    bad = d.abs() > 1e-9 * float(ref.abs().max())
    if not int(bad.sum()):
        continue
    idx = bad.nonzero()
    rows = torch.unique(idx[:, 0]).tolist()
    print(f"rep {rep}: bad {int(bad.sum())} rows {rows[:16]}", flush=True)
    # analyse up to 3 chunks
    done = set()
    for r, col in idx.tolist():
        j = r if r < half else n - 1 - r
        c0 = col - col % 32
        if (j, c0) in done or len(done) >= 4:
            continue
        done.add((j, c0))
        cs = slice(c0, c0 + 32)
        dlo, dhi = d[j, cs], d[n - 1 - j, cs]
        dP, dQ = 0.5 * (dlo + dhi), 0.5 * (dlo - dhi)
        print(f"  chunk pair-row j={j} cols {c0}..{c0+31}: |dP| {float(dP.abs().max()):.3e} |dQ| {float(dQ.abs().max()):.3e}  (tile_m {j//64} wm {(j%64)//32} i {(j%32)//8} g {j%8}; tile_n {c0//128} wn {(c0%128)//32})")
        for name, dd, par in (("P", dP, 0), ("Q", dQ, 1)):
            if float(dd.abs().max()) < 1e-12:
                continue
            ks = torch.arange(par, n, 2, device=dev)               # modes of this parity, folded index k' = position
            Ap, cp = A[j, ks], c[ks][:, cs]                          # [128], [128, 32]
            contrib = torch.stack([Ap[16 * kt:16 * kt + 16] @ cp[16 * kt:16 * kt + 16] for kt in range(kt_n)])   # [kt, 32]
            hit = False
            for kt in range(kt_n):
                for f in (-1.0, 1.0):
                    if float((dd - f * contrib[kt]).abs().max()) < 1e-9:
                        print(f"    {name}: delta = {f:+.0f} x contribution of k-tile {kt}"); hit = True
                for kt2 in range(kt_n):
                    if kt2 == kt: continue
                    # table tile of kt2 used with the data of kt
                    alt = Ap[16 * kt2:16 * kt2 + 16] @ cp[16 * kt:16 * kt + 16]
                    if float((dd - (alt - contrib[kt])).abs().max()) < 1e-9:
                        print(f"    {name}: k-tile {kt} used the TABLE tile of k-tile {kt2}"); hit = True
                    alt = Ap[16 * kt:16 * kt + 16] @ cp[16 * kt2:16 * kt2 + 16]
                    if float((dd - (alt - contrib[kt])).abs().max()) < 1e-9:
                        print(f"    {name}: k-tile {kt} used the DATA tile of k-tile {kt2}"); hit = True
                # other parity's data / other rows' table
                cq = c[torch.arange(1 - par, n, 2, device=dev)][:, cs]
                alt = Ap[16 * kt:16 * kt + 16] @ cq[16 * kt:16 * kt + 16]
                if float((dd - (alt - contrib[kt])).abs().max()) < 1e-9:
                    print(f"    {name}: k-tile {kt} used the data of the OTHER parity"); hit = True
                for j2 in range(half):
                    if j2 == j: continue
                    alt = A[j2, ks][16 * kt:16 * kt + 16] @ cp[16 * kt:16 * kt + 16]
                    if float((dd - (alt - contrib[kt])).abs().max()) < 1e-9:
                        print(f"    {name}: k-tile {kt} used table ROW {j2} (parity {par}) instead of {j}"); hit = True
                    alt = A[j2, torch.arange(1 - par, n, 2, device=dev)][16 * kt:16 * kt + 16] @ cp[16 * kt:16 * kt + 16]
                    if float((dd - (alt - contrib[kt])).abs().max()) < 1e-9:
                        print(f"    {name}: k-tile {kt} used table ROW {j2} of the OTHER parity instead of {j}"); hit = True
            # partial sums in steps of 4 k' (one m8n8k4 step): tail or head of the reduction missing?
            steps = torch.stack([Ap[4 * m_:4 * m_ + 4] @ cp[4 * m_:4 * m_ + 4] for m_ in range(len(Ap) // 4)])     # [32, 32 cols]
            csum = torch.cumsum(steps, 0)
            total = csum[-1]
            for m_ in range(len(Ap) // 4):
                if float((dd - (csum[m_] - total)).abs().max()) < 1e-9:
                    print(f"    {name}: result = partial sum of the first {m_ + 1} of {len(Ap) // 4} k-steps (tail missing)"); hit = True
                if float((dd + csum[m_]).abs().max()) < 1e-9:
                    print(f"    {name}: the first {m_ + 1} k-steps are missing (head missing)"); hit = True
                if float((dd + steps[m_]).abs().max()) < 1e-9:
                    print(f"    {name}: k-step {m_} (k-tile {m_ // 4}, kk {m_ % 4}) is missing"); hit = True
                if float((dd - steps[m_]).abs().max()) < 1e-9:
                    print(f"    {name}: k-step {m_} (k-tile {m_ // 4}, kk {m_ % 4}) was added twice"); hit = True
            if not hit:
                # least squares: which k-tiles explain delta as a combination?
                sol = torch.linalg.lstsq(contrib.T, dd[:, None]).solution[:, 0]
                res = float((contrib.T @ sol - dd).abs().max())
                print(f"    {name}: no single-tile explanation; lstsq over k-tile contributions: coeffs {[round(float(v), 3) for v in sol]} residual {res:.2e}")
    found += 1
    if found >= 3:
        break
print("forensic done, failing reps analysed:", found)
