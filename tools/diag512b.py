import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
dev = torch.device("cuda:0")
n = 512
V = jf.Legendre(n)
rng = np.random.default_rng(0)
Tf = V._dense_table(L.OP_FORWARD, n, n, 0)
u = rng.standard_normal((n, n, n))
ud = torch.from_numpy(u).to(dev)
r0 = np.tensordot(Tf, u, axes=(1, 0))
def report(tag, got):
    d = np.abs(got - r0)
    bad = np.argwhere(d > 1e-9 * np.abs(r0).max())
    print(tag, "max rel", d.max() / np.abs(r0).max(), "n bad", len(bad), flush=True)
    if len(bad):
        print("   k range", bad[:, 0].min(), bad[:, 0].max(), "uniq k", len(np.unique(bad[:, 0])), "| j1 range", bad[:, 1].min(), bad[:, 1].max(),
              "uniq", len(np.unique(bad[:, 1])), "| j2 range", bad[:, 2].min(), bad[:, 2].max(), "uniq", len(np.unique(bad[:, 2])))
        print("   first bad", bad[:5].tolist(), "flat offsets /128:", sorted(set(((b[1] * n + b[2]) // 128) for b in bad[:2000]))[:20])
for rep in range(3):
    report(f"3-D array axis=0 call {rep}", V.forward(ud, axis=0).cpu().numpy())
u2 = ud.reshape(n, n * n)
for rep in range(2):
    report(f"2-D view [512, 262144] call {rep}", V.forward(u2, axis=0).cpu().numpy().reshape(n, n, n))
V2 = jf.Legendre(n)
report("fresh space, 3-D", V2.forward(ud, axis=0).cpu().numpy())
out = torch.empty_like(ud)
p = V2._plans[next(iter(V2._plans))]
p.execute(ud, out); report("same plan into preallocated out", out.cpu().numpy())
torch.cuda.synchronize()
p.execute(ud, out); torch.cuda.synchronize(); report("again after sync", out.cpu().numpy())
