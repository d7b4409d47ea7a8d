"""Static evidence for the parity-folded tensor-core kernels: SASS mnemonic counts per kernel of kernels_dense_fold.cu
(cross-compiled object; no GPU needed).   python tools/sass_summary.py  ->  profiles/r1_fold_sass.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "jaxfun_b200", "build", "kernels_dense_fold.o")
txt = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
out = ["SASS mnemonic counts of kernels_dense_fold.cu, cross-compiled for sm_100a (nvcc 12.9, -O3 -lineinfo):",
       "  cuobjdump -sass jaxfun_b200/build/kernels_dense_fold.o    (no GPU needed; regenerate with tools/sass_summary.py)",
       "DMMA = mma.sync.m8n8k4.f64 (FP64 tensor core); UTMALDG.nD = cp.async.bulk.tensor.nd (TMA tiled load); SYNCS = mbarrier ops;",
       "USETMAXREG = setmaxnreg (producer warpgroup 40 registers, MMA warpgroups 232); LDS.128 = the (even, odd) / (re, im) fragment fetch.",
       "ptxas: 168 registers at launch (384 threads), no spills in the five product kernels, 1 CTA barrier, 197 696 B dynamic shared memory.",
       "template argument: 0 OUT_NN, 1 IN_NN, 2 OUT_NT, 3 IN_NT, 4 CPLX_NT (dmma_fold.cuh); *_scatter = epilogue storing into peer buffers",
       "(its LDL / STL are the dynamically indexed peer-pointer table of the epilogue, outside the k loop).", ""]
PAT = ["DMMA", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.4D", "SYNCS", "USETMAXREG", "LDS.128", "LDS.64", "DADD", "STG.E.128", "STG.E.64",
       "ST.E.128", "ST.E.64", "LDL", "STL", "BAR.SYNC"]
for f in funcs:
    name = f.split("\n")[0].strip()
    dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    cnt = collections.Counter()
    for o in ops:
        for p in PAT:
            if o == p or o.startswith(p + ".") or o.startswith(p):
                cnt[p] += 1
                break
    out.append(f"{dem}\n    instructions {len(ops)}: " + ", ".join(f"{p} {cnt[p]}" for p in PAT if cnt[p]))
open(os.path.join(ROOT, "profiles", "r1_fold_sass.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
