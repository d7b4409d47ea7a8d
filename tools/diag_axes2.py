import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
dev = torch.device("cuda:0")
n, reps = 256, 6
c = torch.randn(n, n, n, dtype=torch.float64, device=dev)
os.environ["JFX_DMMA_FOLD"] = "0"
P = jf.Legendre(n)
refs = {(op, ax): getattr(P, op)(c, axis=ax) for op in ("backward", "forward") for ax in (0, 1)}
os.environ["JFX_DMMA_FOLD"] = "1"
V = jf.Legendre(n)
tot = 0
for (op, ax), ref in refs.items():
    cnt = []
    for r in range(reps):
        u = getattr(V, op)(c, axis=ax); torch.cuda.synchronize()
        cnt.append(int(((u - ref).abs() > 1e-9 * float(ref.abs().max())).sum()))
    tot += sum(cnt)
    print(op, ax, cnt, flush=True)
print("DBG", os.environ.get("JFX_FOLD_DBG", "0"), "total bad", tot)
