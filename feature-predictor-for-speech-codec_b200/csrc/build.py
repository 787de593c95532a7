"""Builds csrc/libfpc_b200.so in-tree with nvcc for sm_100a (no torch involved: the
library is a plain C-ABI shared object, include/fpc_b200.h).

    python feature-predictor-for-speech-codec_b200/csrc/build.py [--force] [--verbose]

-fmad=false: the fp32 path is specified to the rounding; every fused
multiply-add in the kernels is an explicit __fmaf_rn and the compiler must not create
others.  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfpc_b200.so")
SRCS = ["fpc_pack.cu", "fpc_encode_fp32.cu", "fpc_api.cu", "fpc_kmeans.cu", "fpc_kmeans_tc.cu", "fpc_kmeans_ordered.cu", "fpc_umma_selftest.cu", "fpc_encode_bf16.cu", "fpc_ceps2lpc.cu", "fpc_train.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _deps():
    return [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(HERE, "..", "..", "include", "fpc_b200.h")]


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SRCS if os.path.exists(os.path.join(HERE, s))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps()):
        return LIB
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                  "-Xcompiler", "-fPIC", "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
