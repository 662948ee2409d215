#!/bin/bash
# Builds libbgarena.so for sm_100a in-tree (travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${BG_OUT:-../libbgarena.so}
FLAGS="--threads 0 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -cudart static ${BG_NVCC_EXTRA}"
SRCS="api.cu movegen.cu movegen21.cu eval.cu eval128.cu eval_tc.cu select.cu arena.cu two_ply.cu learner.cu hostpath.cu"
$NVCC $FLAGS -shared -o $OUT $SRCS
echo "built $(realpath $OUT)"
