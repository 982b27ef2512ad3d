"""Build kernel-experiment variants of the library: exp/lib_<name>.so = kernels.cu compiled with extra -D flags.
    python tools/exp.py base: plainoff:-DEPI_EXP=1 minb8:-DEPI_MINB=8
(then on the GPU box: EPI_LIB=$PWD/exp/lib_<name>.so python bench.py ...; tools/gpu_ab.sh runs the A/B)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epirust_b200 import build as B
B.build()
os.makedirs(os.path.join(ROOT, "exp"), exist_ok=True)
objs = [os.path.join(B.PKG, "build", s + ".o") for s in B.LIB_SOURCES if s != "kernels.cu"]
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    o = os.path.join(ROOT, "exp", f"kernels_{name}.o")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include")] + defs.split() + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "kernels.cu"), "-o", o],
                          stderr=open(os.path.join(ROOT, "exp", f"ptxas_{name}.log"), "w"))
    subprocess.check_call([B._nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp", "-o", os.path.join(ROOT, "exp", f"lib_{name}.so"), o] + objs + ["-lnccl"])
    print("built", name, defs)
