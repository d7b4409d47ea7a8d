import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = torch.randn(n, n, n, dtype=torch.float64, device=dev)
os.environ["JFX_DMMA_FOLD"] = "0"
P = jf.Legendre(n)
refs = {(op, ax): getattr(P, op)(c, axis=ax) for op in ("backward", "forward") for ax in range(3)}
os.environ["JFX_DMMA_FOLD"] = "1"
V = jf.Legendre(n)
for (op, ax), ref in refs.items():
    for r in range(reps):
        u = getattr(V, op)(c, axis=ax); torch.cuda.synchronize()
        d = (u - ref).abs()
        bad = d > 1e-9 * float(ref.abs().max())
        nb = int(bad.sum())
        msg = f"{op} axis {ax} rep {r}: bad {nb}"
        if nb:
            idx = bad.nonzero()
            msg += " | " + " ; ".join(f"dim{k}: {len(torch.unique(idx[:, k]))} uniq, first {torch.unique(idx[:, k])[:8].tolist()}" for k in range(3))
        print(msg, flush=True)
