"""cuBLAS DGEMM on the shapes of the per-axis contraction (calibration only, not on the product path)."""
import torch
dev = torch.device("cuda:0")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = 256
X = torch.randn(n * n, n, dtype=torch.float64, device=dev)
T = torch.randn(n, n, dtype=torch.float64, device=dev)
out = torch.empty_like(X)
ms = t(lambda: torch.matmul(X, T.t(), out=out))
print(f"last axis  X[65536,256] @ T^T : {ms*1e3:7.1f} us  {2*n**4/ms/1e9:6.2f} TFLOP/s")
X3 = X.view(n, n, n)
out3 = torch.empty_like(X3)
ms = t(lambda: torch.matmul(T, X3, out=out3))
print(f"middle axis T @ X[o] batched  : {ms*1e3:7.1f} us  {2*n**4/ms/1e9:6.2f} TFLOP/s")
X0 = X.view(n, n * n)
out0 = torch.empty_like(X0)
ms = t(lambda: torch.matmul(T, X0, out=out0))
print(f"first axis T @ X[256,65536]   : {ms*1e3:7.1f} us  {2*n**4/ms/1e9:6.2f} TFLOP/s")
A = torch.randn(8192, 8192, dtype=torch.float64, device=dev); B = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
ms = t(lambda: torch.matmul(A, B), reps=3)
print(f"8192^3: {2*8192**3/ms/1e9:6.2f} TFLOP/s")
