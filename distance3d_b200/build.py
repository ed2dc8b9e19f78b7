"""Builds libd3d_b200.so in-tree with nvcc for sm_100a.

-fmad=false is part of the numerical contract: the reference's scalar code
never fuses multiply-adds, the BLAS FMA patterns are written out explicitly in
csrc/d3d_math.cuh.
"""
import concurrent.futures
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libd3d_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        # *_f32.cu only #include their fp64 twin: every .cu is a dependency of theirs
        deps = [src] + headers + (sources if src.endswith("_f32.cu") else [])
        if force or _stale(obj, deps):
            flags = list(NVCC_FLAGS)
            if src.endswith("_f32.cu"):
                flags.remove("-fmad=false")  # the fp32 mode has no bit-level contract
            r = subprocess.run([nvcc] + flags + ["-c", src, "-o", obj],
                               capture_output=True, text=True)
            logs[src] = r.stdout + r.stderr
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, logs[src]))
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(OUT, objs):
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        for src, log in logs.items():
            print(log)
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        for src, log in logs.items():
            f.write("==== %s\n%s\n" % (src, log))
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
