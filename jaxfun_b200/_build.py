"""Build the in-tree CUDA library (libjfx.so) for sm_100a with nvcc.

The shared object is placed next to the package (jaxfun_b200/libjfx.so) so it travels with the
repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjfx.so")
SOURCES = ["jfx_api.cu", "kernels_dense.cu", "kernels_dense_tma.cu", "kernels_dense_fold.cu", "kernels_fft.cu", "kernels_fft2.cu", "kernels_fft2_pair.cu", "kernels_fft2_stream.cu", "kernels_fused.cu", "kernels_pointwise.cu", "kernels_banded.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libjfx.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "jfx.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _read_deps(path: str):
    """Repository files named by an nvcc -MD dependency file (system / CUDA headers are ignored); [] if unreadable."""
    try:
        txt = open(path).read()
    except OSError:
        return []
    root = os.path.abspath(os.path.join(HERE, ".."))
    out = []
    for tok in txt.split():
        tok = tok.rstrip(":")
        if tok.startswith(root) and os.path.exists(tok) and not tok.endswith(".o"):
            out.append(tok)
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [
        os.path.join(HERE, "..", "include", "jfx.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    extra = os.environ.get("JFX_NVCC_EXTRA", "")
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        # incremental: an object newer than its source and the headers it includes (nvcc -MD dependency file; without one:
        # every header) is kept (force / verbose / extra flags rebuild all)
        if not force and not verbose and not extra and os.path.exists(obj):
            deps = _read_deps(obj + ".d")
            newest = max(os.path.getmtime(d) for d in deps) if deps else newest_header
            if os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(CSRC, src)), newest):
                continue
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("JFX_NVCC_EXTRA", "").split(), "-MD", "-MF", obj + ".d",
               "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    if not procs and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(o) for o in objs):
        return LIB                       # every object is newer than its sources and the library newer than every object
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    tmp = LIB + ".tmp"
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-cudart", "static"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
