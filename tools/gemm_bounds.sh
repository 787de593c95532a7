#!/bin/bash
# Debug builds of the fp32 frame kernel that remove one ingredient of the gate GEMM at a time (the weight stream,
# or the arithmetic) (results are garbage,
# only the GRU phase timer of tools/phase_profile.py is meaningful).  Usage: tools/gemm_bounds.sh  (then run on the GPU:
#   for l in csrc/dbg/*.so; do FPC_B200_LIB=$l python tools/phase_profile.py 4096 100 | head -2; done)
set -e
cd "$(dirname "$0")/../feature-predictor-for-speech-codec_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OTHERS="fpc_pack.o fpc_api.o fpc_kmeans.o fpc_umma_selftest.o fpc_encode_bf16.o fpc_ceps2lpc.o fpc_train.o"
mkdir -p dbg
build() { # name, defines
  nvcc $F $2 -c fpc_encode_fp32.cu -o /tmp/enc_$1.o && nvcc -shared -o dbg/libfpc_$1.so /tmp/enc_$1.o $OTHERS -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
}
build stream_only "-DFPC_DEBUG_STREAM_ONLY" &
build no_stream "-DFPC_DEBUG_NO_STREAM" &
wait
ls -la dbg
