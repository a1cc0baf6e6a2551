"""Builds libspllt_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The shared library is the product: everything under tests/ and bench.py calls it through
ctypes (spllt_b200/api.py).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libspllt_b200.so")
SOURCES = ["kernels.cu", "solve_pipe.cu", "engine.cu", "capi.cu", "analyse.cpp", "symbolic.cpp"]
HEADERS = ["kernels.cuh", "engine.h", "model.h", "symbolic.h",
           os.path.join("..", "..", "include", "spllt_iface.h"),
           os.path.join("..", "..", "include", "spllt_b200.h")]
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
METIS = os.path.join(CUDA, "targets", "x86_64-linux", "lib", "libmetis_static.a")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.path.join(CUDA, "bin", "nvcc")
    objs = []
    bdir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(bdir, exist_ok=True)
    common = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fopenmp", "-I", CSRC]
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s + ".o")
        cmd = [nvcc] + common + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + [METIS, "-cudart", "static", "-Xcompiler", "-fopenmp", "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
