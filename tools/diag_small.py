import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
os.environ["JFX_DMMA_FOLD"] = "0"
P = jf.TensorProduct(*[jf.Legendre(n)] * 3)
c = torch.randn(n, n, n, dtype=torch.float64, device=dev)
ref = P.backward(c)
os.environ["JFX_DMMA_FOLD"] = "1"
T = jf.TensorProduct(*[jf.Legendre(n)] * 3)
for r in range(reps):
    u = T.backward(c); torch.cuda.synchronize()
    d = (u - ref).abs()
    bad = d > 1e-9 * float(ref.abs().max())
    print("rep", r, "bad", int(bad.sum()), "max rel", float(d.max() / ref.abs().max()), flush=True)
    if int(bad.sum()):
        idx = bad.nonzero()
        for ax in range(3):
            print("   axis", ax, "uniq", len(torch.unique(idx[:, ax])), "first", torch.unique(idx[:, ax])[:12].tolist())
