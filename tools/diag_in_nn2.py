import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
dev = torch.device("cuda:0")
n, inner = 256, 1 << 18
os.environ["JFX_DMMA_FOLD"] = "0"
P = jf.Legendre(n)
u = torch.randn(n, inner, dtype=torch.float64, device=dev)
ref = P.forward(u, axis=0)
os.environ["JFX_DMMA_FOLD"] = "1"
Tf = torch.from_numpy(P._dense_table(L.OP_FORWARD, n, n, 0)).to(dev)      # [k, j]
half = n // 2
for call in range(6):
    V = jf.Legendre(n)
    got = V.forward(u, axis=0); torch.cuda.synchronize()
    d = got - ref
    bad = (d.abs() > 1e-9 * float(ref.abs().max()))
    nb = int(bad.sum())
    print("call", call, "bad", nb, flush=True)
    if not nb:
        continue
    idx = bad.nonzero()
    rows = torch.unique(idx[:, 0])
    print("  rows (first 20):", rows[:20].tolist(), " n rows", len(rows), " parity of rows:", set((rows % 2).tolist()))
    # chunk structure
    r0 = int(idx[0, 0]); cols = idx[idx[:, 0] == r0][:, 1]
    c0 = int(cols.min()); print("  row", r0, "bad cols", c0, "..", int(cols.max()), "count", len(cols), "tile_n", c0 // 128, "wn", (c0 % 128) // 32)
    # which CTA / tile order: tl = tm * tiles_n? (tn fastest): tl = tm*tiles_n + tn ; CTA = tl % 148 ; iteration = tl // 148
    tiles_n = inner // 128
    tls = set()
    for r in rows.tolist()[:2000]:
        kidx = r // 2
        tm = kidx // 64
        for c in torch.unique(idx[idx[:, 0] == r][:, 1] // 128).tolist():
            tls.add(tm * tiles_n + c)
    tls = sorted(tls)
    print("  tiles hit:", len(tls), "iteration index (tl // 148) histogram:", np.bincount(np.array(tls) // 148)[:40].tolist())
    print("  CTA ids (tl % 148) of first 30:", [t % 148 for t in tls[:30]])
    # error decomposition for the first bad chunk: contribution of each k-tile (16 folded j's) to row r0 on these 32 columns
    cs = slice(c0 - c0 % 32, c0 - c0 % 32 + 32)
    sig = 1.0 if r0 % 2 == 0 else -1.0
    um = u[:half, cs] + sig * u.flip(0)[:half, cs]                       # u_j + sigma u_{n-1-j}
    contrib = torch.stack([Tf[r0, kt * 16:(kt + 1) * 16] @ um[kt * 16:(kt + 1) * 16] for kt in range(half // 16)])   # [kts, 32]
    delta = d[r0, cs]
    print("  |delta| max", float(delta.abs().max()), " |ref| max", float(ref[r0, cs].abs().max()))
    for kt in range(half // 16):
        for f in (-1.0, -2.0, 1.0):
            if float((delta - f * contrib[kt]).abs().max()) < 1e-9:
                print(f"  delta == {f} x contribution of k-tile {kt}")
    # does delta match using the PLUS combination (u + u') instead of minus for some k-tile?
    up = u[:half, cs] - sig * u.flip(0)[:half, cs]
    for kt in range(half // 16):
        alt = Tf[r0, kt * 16:(kt + 1) * 16] @ up[kt * 16:(kt + 1) * 16]
        if float((delta - (alt - contrib[kt])).abs().max()) < 1e-9:
            print(f"  k-tile {kt}: used the opposite sign combination")
    # a neighbouring row's table entries?
    for dr in (-2, -1, 1, 2, 16, -16):
        rr = r0 + dr
        if 0 <= rr < n:
            for kt in range(half // 16):
                alt = Tf[rr, kt * 16:(kt + 1) * 16] @ um[kt * 16:(kt + 1) * 16]
                if float((delta - (alt - contrib[kt])).abs().max()) < 1e-9:
                    print(f"  k-tile {kt}: table row {rr} used instead of {r0}")
