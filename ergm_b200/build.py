"""Builds libergm_b200.so (all CUDA kernels + the C ABI) in-tree for sm_100a.

    python -m ergm_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the
gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libergm_b200.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("ERGM_NVCC_EXTRA", "").split()   # e.g. -DERGM_ATTN_TRACE for scripts/trace_attn_bwd.py (never shipped)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/ergm_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _hdr_digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/ergm_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and not f.endswith(".cu"):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libergm_b200.stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdig = _hdr_digest()

    def compile_one(src):
        srcp = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        ostamp = obj + ".stamp"
        key = hashlib.sha256(open(srcp, "rb").read() + hdig.encode()).hexdigest()
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == key:
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + ["-c", srcp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        open(ostamp, "w").write(key)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = list(ex.map(compile_one, _sources()))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        open(os.path.join(OBJDIR, "ptxas.log"), "a").write(log)
        if verbose:
            print(log)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
