import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else None
def run(n, inner, reps=4):
    os.environ["JFX_DMMA_FOLD"] = "0"
    P = jf.Legendre(n)
    u = torch.randn(n, inner, dtype=torch.float64, device=dev)
    ref = P.forward(u, axis=0)
    os.environ["JFX_DMMA_FOLD"] = "1"
    V = jf.Legendre(n)
    res = []
    for _ in range(reps):
        got = V.forward(u, axis=0); torch.cuda.synchronize()
        d = (got - ref).abs()
        bad = (d > 1e-9 * float(ref.abs().max()))
        rows = torch.unique(bad.nonzero()[:, 0]).tolist() if int(bad.sum()) else []
        res.append((int(bad.sum()), rows[:6]))
    print(f"n={n} inner={inner}: bad counts/rows per call {res}", flush=True)
if only:
    n, inner = (int(v) for v in only.split("x"))
    run(n, inner, 2)
else:
    for n, inner in ((512, 4096), (512, 32768), (512, 65536), (512, 131072), (512, 262144), (256, 262144), (256, 1 << 20), (384, 262144), (1024, 65536), (1024, 131072)):
        run(n, inner)
