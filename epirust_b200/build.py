"""In-tree build of the sm_100a shared library (and the engine-app CLI) with nvcc.

    python -m epirust_b200.build [--force]

Produces epirust_b200/libepirust_b200.so (the C ABI of include/epi.h) and epirust_b200/engine-app.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libepirust_b200.so")
APP = os.path.join(PKG, "engine-app")

LIB_SOURCES = ["kernels.cu", "tiles.cu", "travel.cu", "engine.cpp", "host_model.cpp", "json.cpp", "simulation.cpp", "travel.cpp", "multi.cpp", "configuration.cpp"]
APP_SOURCES = ["engine_app_main.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-fopenmp",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = list(sources) + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "epi.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force or _stale(LIB, srcs):
        objs = []
        obj_dir = os.path.join(PKG, "build")
        os.makedirs(obj_dir, exist_ok=True)
        for s in srcs:
            o = os.path.join(obj_dir, os.path.basename(s) + ".o")
            if force or _stale(o, [s]):
                cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o]
                if s.endswith(".cu"):
                    cmd += ["-Xptxas", "-v"] if verbose else []
                if verbose:
                    print(" ".join(cmd))
                subprocess.check_call(cmd)
            objs.append(o)
        # libnccl: the traveller exchange of multi-region runs (csrc/multi.cpp)
        cmd = [nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp", "-o", LIB] + objs + ["-lnccl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    app_srcs = [os.path.join(CSRC, s) for s in APP_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if app_srcs and (force or _stale(APP, app_srcs + [LIB])):
        cmd = [nvcc, "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include")] + app_srcs + [
            "-o", APP, "-L", PKG, "-lepirust_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
