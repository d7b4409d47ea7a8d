"""Regenerate the tracked summaries under profiles/ from the scratch captures in gpurun_out/.

    python tools/make_profiles.py r1d      # uses gpurun_out/{launches,prof_dgemm256,prof_fft2_cheb256,bench}_<tag>.*
"""
import collections, csv, io, json, re, subprocess, sys, os

tag = sys.argv[1] if len(sys.argv) > 1 else "r1d"
RD = sys.argv[2] if len(sys.argv) > 2 else "r1"          # prefix of the files written under profiles/ (r1, r2, ...)
# JFX_PROF_IN / JFX_PROF_OUT: run on the GPU box right after the captures (the .ncu-rep files are too large to travel back)
G, P = os.environ.get("JFX_PROF_IN", "gpurun_out"), os.environ.get("JFX_PROF_OUT", "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------------
rows = list(csv.reader(open(f"{G}/launches_{tag}.csv")))
hdr, per = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("void ", "")
    if "at::" in name:
        short = "torch (input generation)"
    elif "dgemm_dmma" not in name and ("cutlass" in name or "gemm" in name.lower() or "cublas" in name.lower()):
        short = "cuBLAS DGEMM (FP64 peak calibration)"
    elif "dgemm_dmma_tma" in name and any("dgemm_dmma_fold" in r2[hdr.index("Kernel Name")] for r2 in rows if hdr and len(r2) >= len(hdr) and r2[0] != "ID"):
        short = short + "  [fold_ab: the plain kernel of the live A/B, outside the timed step]"
    d = per.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += float(r[hdr.index("Metric Value")]) / 1e3
step = {k: v for k, v in per.items() if ("dgemm_dmma" in k or "fft2_kernel" in k) and "fold_ab" not in k}
tot = sum(v[1] for v in step.values())
bench = json.loads(open(f"{G}/bench_{tag}.json").read().splitlines()[-1])
dl, dc = bench["detail"]["legendre3"]["ms_per_pair"], bench["detail"]["chebyshev3"]["ms_per_pair"]
out = ["ncu launch list of: python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e   (gpurun, 1x B200, --clock-control none)",
       "per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes", ""]
for k, (n, t) in per.items():
    share = f"{100 * t / tot:5.1f}% of the step's kernel time" if k in step else "(outside the timed step)"
    out.append(f"{k:72s} launches={n:4d} total={t:10.1f} us avg={t / n:8.1f} us  {share}")
dg = sum(v[1] for k, v in step.items() if "dgemm" in k)
out += ["", f"dgemm_dmma share of the step under ncu: {100 * dg / tot:.1f}%   |   live CUDA-event split of the same step "
            f"(bench_{tag}.json): legendre3 {dl:.3f} ms = {100 * dl / (dl + dc):.1f}%, chebyshev3 {dc:.3f} ms = {100 * dc / (dl + dc):.1f}%"]
open(f"{P}/{RD}_launches_summary.txt", "w").write("\n".join(out) + "\n")
open(f"{P}/{RD}_launches.csv", "w").write(open(f"{G}/launches_{tag}.csv").read())

# ---- full captures ---------------------------------------------------------------------------------------
def summary(rep, header):
    o = subprocess.run([sys.executable, "tools/ncu_summary.py", rep, "--ops"], capture_output=True, text=True).stdout
    return header + "\n" + o

DG = f"{G}/prof_dgemm256_{tag}.ncu-rep" if os.path.exists(f"{G}/prof_dgemm256_{tag}.ncu-rep") else f"{G}/prof_fold256_{tag}.ncu-rep"
open(f"{P}/{RD}_dgemm_dmma_ncu.txt", "w").write(summary(
    DG, "ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6  python tools/profile_step.py legendre 256\n"
        "(six launches = the axis passes of one Legendre^3 256^3 backward (variants <0>, <0>, <2>) and forward (<1>, <1>, <3>))"))
open(f"{P}/{RD}_fft2_cheb256_ncu.txt", "w").write(summary(
    f"{G}/prof_fft2_cheb256_{tag}.ncu-rep",
    "ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6  python tools/profile_step.py chebyshev 256"))

def traffic(path, pat):
    o = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(o)))
    h, u = rows[0], rows[1]
    ir, iw, ik = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Kernel Name")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    return [float(r[ir]) * scale[u[ir]] + float(r[iw]) * scale[u[iw]] for r in rows[2:] if pat in r[ik]]

t = {}
v = traffic(DG, "dgemm_dmma")
t["dgemm_dmma"] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v),
                   "algorithmic_bytes_per_launch": 2 * 8 * 256**3 + 8 * 256 * 256,
                   "source": f"profiles/{RD}_dgemm_dmma_ncu.txt (ncu --set full; the axis passes of Legendre^3 256^3 transforms)"}
vf = traffic(DG, "dgemm_dmma_fold")
if vf:   # the parity-folded kernel (default since the end of round 1): what bench.py looks up when the step is folded
    t["dgemm_dmma_fold"] = {"dram_bytes_per_launch": sum(vf) / len(vf), "launches": len(vf),
                            "algorithmic_bytes_per_launch": 2 * 8 * 256**3 + 8 * 256 * 256,
                            "source": f"profiles/{RD}_dgemm_dmma_ncu.txt (ncu --set full; the axis passes of Legendre^3 256^3 transforms)"}
v = traffic(f"{G}/prof_fft2_cheb256_{tag}.ncu-rep", "fft2_kernel")
t["fft2_kernel"] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v), "algorithmic_bytes_per_launch": 2 * 8 * 256**3,
                    "source": f"profiles/{RD}_fft2_cheb256_ncu.txt (ncu --set full; 6 launches = backward + forward Chebyshev^3 256^3); "
                              "writes still dirty in L2 at kernel end are not counted by dram__bytes_write"}
json.dump(t, open(f"{P}/traffic_{RD}.json", "w"), indent=1)
open(f"{P}/{RD}_bench.json", "w").write(json.dumps(bench, indent=1) + "\n")
print("\n".join(out))
print(json.dumps(t, indent=1))
