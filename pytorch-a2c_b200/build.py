"""Builds liba2c_b200.so (sm_100a) in-tree with nvcc.  No torch involved: the
library is a plain C-ABI shared object (include/a2c_b200.h).

    python pytorch-a2c_b200/build.py [--force]

Profiling builds (never the shipped library): A2C_NVCC_FLAGS="-DA2C_DEBUG_KNOBS" enables the stage-isolation
environment knobs, A2C_NVCC_FLAGS="-DA2C_TC_TRACE" the per-role clock stamps of the tcgen05 GEMM and the
weight-gradient kernel (a2c_debug_tc_trace / a2c_debug_wg_trace, DESIGN.md lessons 19 and 22).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "a2c_b200", "lib")
LIB = os.path.join(OUT_DIR, "liba2c_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-DNDEBUG"] + os.environ.get("A2C_NVCC_FLAGS", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/a2c_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp_file = os.path.join(OUT_DIR, "build.stamp")
    stamp = _stamp()
    if (not force and os.path.isfile(LIB) and os.path.isfile(stamp_file)
            and open(stamp_file).read() == stamp):
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
