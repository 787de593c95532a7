"""Build recipe for the CPU oracle (test infrastructure only).

    python oracle/build.py        ->  oracle/_build/libfpc_oracle.so

-ffp-contract=off is mandatory (every fused multiply-add in the oracle is an explicit
fmaf/fma call; the compiler must not invent others).  -mfma/-mavx2 make fmaf a single
instruction; the library is built in the build container and travels to the GPU box,
so no -march=native.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfpc_oracle.so")
SRCS = [os.path.join(HERE, "fpc_oracle.c")]
DEPS = SRCS + [os.path.join(HERE, "fpc_oracle_vq.inc")]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and all(
            os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2",
           "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-o", LIB] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
