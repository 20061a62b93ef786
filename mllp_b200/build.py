"""Builds mllp_b200/csrc/libmllp_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo
snapshot to the GPU box.  No torch involvement: the library exposes a plain C ABI
(include/mllp_b200.h) and is loaded with ctypes (mllp_b200/_cabi.py).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(CSRC, "libmllp_b200.so")
SOURCES = ["lp_format.cpp", "pdhg_kernels.cu", "batch_kernels.cu", "gnn_kernels.cu", "gnn_backward.cu", "blocks.cu", "scaling.cu", "multicast.cu", "cabi.cu"]
HEADERS = ["lp_format.h", "pdhg_kernels.cuh", "pdhg_host.h", "gnn_common.cuh", os.path.join(ROOT, "include", "mllp_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    return None


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = SO + ".stamp"
    digest = _digest()
    if not force and os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read() == digest:
        return SO
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    common = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]
    cxx = _host_cxx()
    if cxx:
        common += ["-ccbin", cxx]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("MLLP_NVCC_FLAGS", "").split()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, *ARCH, "-shared", "-o", SO, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    if cxx:
        link += ["-ccbin", cxx]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(digest)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
