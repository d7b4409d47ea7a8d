"""Where does the slab transform's time go?  Per rank, CUDA events: phase-1 passes without the scatter, the scatter pass
alone, barrier, phase 2, whole transform; host time per transform.  torchrun --nproc-per-node P tools/slab_breakdown.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L, sharding as S

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = jf.TensorProduct(*[jf.Legendre(n)] * 3)
D = S.SlabTensorProduct(T)
c = torch.randn(n // world, n, n, dtype=torch.float64, device=dev)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, 1e3 * th / reps

be = D._backend(L.OP_BACKWARD)
res = {}
res["whole backward (slab)"] = timed(lambda: D.backward(c))
u = D.backward(c)
res["whole forward (slab)"] = timed(lambda: D.forward(u))
# pieces of backward: phase 1 local (axes 1, 2) with ordinary stores, phase 2 (axis 0) on the exchanged shape
res["phase 1 local passes, plain stores"] = timed(lambda: be.apply_axes(c, [1, 2]))
y = torch.randn(n, n // world, n, dtype=torch.float64, device=dev)
res["phase 2 (axis 0)"] = timed(lambda: be.apply_axes(y, [0]))
# phase 1 with the scatter epilogue but no barrier (peer stores still go over NVLink)
ent = [v for v in be._plans.values() if isinstance(v, dict) and "hdls" in v][0]
plan = be._plan_for(c, [1, 2])
hdl = ent["hdls"][0]
ptrs = [int(p) for p in hdl.buffer_ptrs]
res["phase 1 with scatter epilogue, no barrier"] = timed(lambda: plan.execute_scatter(c, ptrs, rank, 1))
res["barrier alone"] = timed(lambda: hdl.barrier(channel=0))
# pieces of forward: phase 1 = axes 0, 2 of the physical block [n, n/P, n]; phase 2 = axis 1 of [n/P, n, n]
bf = D._backend(L.OP_FORWARD)
res["fwd phase 1 local passes (axes 0, 2), plain stores"] = timed(lambda: bf.apply_axes(u, [0, 2]))
res["fwd phase 1 axis 0 only"] = timed(lambda: bf.apply_axes(u, [0]))
res["fwd phase 1 axis 2 only"] = timed(lambda: bf.apply_axes(u, [2]))
z = torch.randn(n // world, n, n, dtype=torch.float64, device=dev)
res["fwd phase 2 (axis 1)"] = timed(lambda: bf.apply_axes(z, [1]))
entf = [v for v in bf._plans.values() if isinstance(v, dict) and "hdls" in v][0]
planf = bf._plan_for(u, [0, 2])
ptrsf = [int(p) for p in entf["hdls"][0].buffer_ptrs]
res["fwd phase 1 with scatter epilogue, no barrier"] = timed(lambda: planf.execute_scatter(u, ptrsf, rank, 0))
if rank == 0:
    for k, (g, h) in res.items():
        print(f"P={world} {k:45s} gpu {g:7.3f} ms   host {h:7.3f} ms", flush=True)
dist.destroy_process_group()
