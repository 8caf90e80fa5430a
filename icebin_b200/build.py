"""Compile the CUDA sources in icebin_b200/csrc into icebin_b200/libicebin_b200.so (sm_100a, in-tree).

    python -m icebin_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  build.cu and apply.cu are compiled with -fmad=false so that matrix values and
un-split SpMM rows are bit-identical to the CPU oracle's mul-then-add arithmetic.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libicebin_b200.so")
OBJ = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"]
UNITS = {
    "build.cu": ["-fmad=false"],
    "apply.cu": ["-fmad=false"],
    "apply_stream.cu": ["-fmad=false"],
    "apply_group.cu": ["-fmad=false"],
    "smooth.cu": ["-fmad=false"],
    "couple.cu": ["-fmad=false"],
    "exgrid.cu": ["-fmad=false"],
    "hntr.cu": [],
    "merge.cu": ["-fmad=false"],
    "xchg.cu": [],
    "modele.cu": ["-fmad=false"],
    "capi.cu": [],
}


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    raise RuntimeError("nvcc not found")


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False, defines=(), out=None, objdir=None):
    """defines/out/objdir: instrumented variants for tools/ (e.g. -DIB_STREAM_TRACE); the product build uses none."""
    global OUT, OBJ
    saved = (OUT, OBJ)
    if out:
        OUT, OBJ = out, objdir or (out + ".build")
    try:
        return _build(force, verbose, ["-D" + d for d in defines])
    finally:
        OUT, OBJ = saved


def _build(force, verbose, extra):
    dep = _newest_dep()
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= dep:
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    units = {u: f for u, f in UNITS.items() if os.path.exists(os.path.join(CSRC, u))}

    def cc(item):
        src, flags = item
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep:
            return obj
        cmd = [nvcc] + ARCH + COMMON + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on " + src)
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        objs = list(ex.map(cc, units.items()))
    link = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
