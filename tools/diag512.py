"""Diagnostic: which pass of the 512^3 Legendre forward is wrong?  (GPU; compares single-axis applications with numpy matmuls)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
V = jf.Legendre(n)
rng = np.random.default_rng(0)
Tf = V._dense_table(L.OP_FORWARD, n, n, 0)
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
# 1-D batched (IN_NT)
for rows in (100, 4096, 262144):
    u = rng.standard_normal((rows, n))
    got = V.forward(torch.from_numpy(u).to(dev)).cpu().numpy()
    print("IN_NT rows", rows, rel(got, u @ Tf.T), flush=True)
# axis 0 (IN_NN) with various inner
for inner in (2, 256, 512, 4096, 262144):
    u = rng.standard_normal((n, inner))
    got = V.forward(torch.from_numpy(u).to(dev), axis=0).cpu().numpy()
    print("IN_NN inner", inner, rel(got, Tf @ u), flush=True)
# middle axis
for outer, inner in ((8, 512), (512, 512), (64, 64)):
    u = rng.standard_normal((outer, n, inner))
    got = V.forward(torch.from_numpy(u).to(dev), axis=1).cpu().numpy()
    ref = np.einsum("kj,oji->oki", Tf, u)
    print("IN_NN middle outer", outer, "inner", inner, rel(got, ref), flush=True)
# the 3-D plan pass by pass
T3 = jf.TensorProduct(V, V, V)
u = rng.standard_normal((n, n, n))
ud = torch.from_numpy(u).to(dev)
full = T3.forward(ud).cpu().numpy()
a0 = V.forward(ud, axis=0); a1 = V.forward(a0, axis=1); a2 = V.forward(a1, axis=2)
print("3-D plan vs three single-axis calls (GPU both)", rel(full, a2.cpu().numpy()), flush=True)
r0 = np.tensordot(Tf, u, axes=(1, 0))
print("axis0 vs numpy", rel(a0.cpu().numpy(), r0), flush=True)
r1 = np.einsum("kj,oji->oki", Tf, r0)
print("axis1 vs numpy", rel(a1.cpu().numpy(), r1), flush=True)
r2 = r1 @ Tf.T
print("axis2 vs numpy", rel(a2.cpu().numpy(), r2), "full vs numpy", rel(full, r2), flush=True)
os.environ["JFX_DMMA_FOLD"] = "0"
T3p = jf.TensorProduct(jf.Legendre(n), jf.Legendre(n), jf.Legendre(n))
print("plain kernel full vs numpy", rel(T3p.forward(ud).cpu().numpy(), r2), flush=True)
