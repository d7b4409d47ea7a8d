"""Build a variant of libjfx.so with extra nvcc flags (macro switches) for A/B timing on the GPU box.

    python tools/build_variant.py TAG -DJFX_PLAN256_884 -DJFX_FFT_REGS8=64
    JFX_LIB_PATH=jaxfun_b200/variants/libjfx_TAG.so python tools/bench_axes.py cheb

Objects go to jaxfun_b200/build_TAG/, the library to jaxfun_b200/variants/ (git-ignored, travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("jfx_build", os.path.join(HERE, "jaxfun_b200", "_build.py"))
B = importlib.util.module_from_spec(_spec)      # the build script alone: importing the package would load libjfx.so
_spec.loader.exec_module(B)


def main():
    tag, flags = sys.argv[1], sys.argv[2:]
    only = None                      # --only=a.cu,b.cu: recompile these, link the rest from the main build directory
    for f in list(flags):
        if f.startswith("--only="):
            only = f[len("--only="):].split(",")
            flags.remove(f)
    bdir = os.path.join(B.HERE, "build_" + tag)
    odir = os.path.join(B.HERE, "variants")
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(odir, exist_ok=True)
    nvcc = B._nvcc()
    procs, objs = [], []
    for src in B.SOURCES:
        if only is not None and src not in only:
            objs.append(os.path.join(B.HERE, "build", src.replace(".cu", ".o")))
            continue
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise SystemExit(f"nvcc failed on {src}")
    lib = os.path.join(odir, f"libjfx_{tag}.so")
    subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-cudart", "static"])
    print(lib)


if __name__ == "__main__":
    main()
