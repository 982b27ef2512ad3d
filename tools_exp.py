"""Build kernel-experiment variants of the library: exp/lib_<n>.so = kernels.cu compiled with -DEPI_EXP=<n>.
    python tools_exp.py 0 1 2 3      (then on the GPU box: EPI_LIB=exp/lib_1.so python bench.py ...)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from epirust_b200 import build as B
B.build()
os.makedirs(os.path.join(ROOT, "exp"), exist_ok=True)
objs = [os.path.join(B.PKG, "build", s + ".o") for s in B.LIB_SOURCES if s != "kernels.cu"]
for n in sys.argv[1:]:  # "3" or "3:8" = EPI_EXP 3 with EPI_PF 8
    o = os.path.join(ROOT, "exp", f"kernels_{n.replace(':', '_')}.o")
    exp, _, pf = n.partition(":")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), f"-DEPI_EXP={exp}"] + ([f"-DEPI_PF={pf}"] if pf else []) + [ "-Xptxas", "-v", "-c", os.path.join(B.CSRC, "kernels.cu"), "-o", o],
                          stderr=open(os.path.join(ROOT, "exp", f"ptxas_{n.replace(':', '_')}.log"), "w"))
    subprocess.check_call([B._nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp", "-o", os.path.join(ROOT, "exp", f"lib_{n.replace(':', '_')}.so"), o] + objs)
    print("built", n)
