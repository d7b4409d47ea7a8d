"""First execution of freshly created plans vs settled executions (bitwise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ud = torch.randn(n, n, n, dtype=torch.float64, device=dev)
V = jf.Legendre(n)
kinds = {"fwd0": lambda S, o=None: S.forward(ud, axis=0), "bwd0": lambda S: S.backward(ud, axis=0), "fwd2": lambda S: S.forward(ud, axis=2),
         "fwd1": lambda S: S.forward(ud, axis=1), "bwd2": lambda S: S.backward(ud, axis=2)}
good = {}
for k, fn in kinds.items():
    for _ in range(3):
        good[k] = fn(V).clone()
torch.cuda.synchronize()
# settled plan, output prefilled with NaN: unwritten elements would show
for k in kinds:
    plan = [p for key, p in V._plans.items()][list(kinds).index(k)]
    out = torch.full_like(ud, float("nan"))
    plan.execute(ud, out); torch.cuda.synchronize()
    print(k, "settled plan into NaN-filled out: nan count", int(torch.isnan(out).sum()), "diff vs good", int((out != good[k]).sum()), flush=True)
tot = {k: 0 for k in kinds}
for i in range(reps):
    W = jf.Legendre(n)
    for k, fn in kinds.items():
        a = fn(W); torch.cuda.synchronize()
        b = fn(W); torch.cuda.synchronize()
        d1, d2 = (a != good[k]), (b != good[k])
        n1, n2 = int(d1.sum()), int(d2.sum())
        if n1 or n2:
            tot[k] += 1
            idx = d1.nonzero()[:3].tolist() if n1 else d2.nonzero()[:3].tolist()
            mx = float((a - good[k]).abs().max()), float((b - good[k]).abs().max())
            print(f"rep {i} {k}: first call {n1} differing (max abs {mx[0]:.3e}), second call {n2} (max abs {mx[1]:.3e}) at {idx}", flush=True)
print("STRESS", n, tot)
