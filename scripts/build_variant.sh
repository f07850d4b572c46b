#!/bin/bash
# Build a variant of libsgs.so for same-box A/B runs: scripts/build_variant.sh NAME "-DSGS_DFS_MINB=8 ..."
# → build_ab/libsgs_NAME.so (only sparse_search.cu is recompiled with the extra flags)
set -e
cd "$(dirname "$0")/../stabilizer_gs_b200/csrc"
NAME=$1; FLAGS=$2
mkdir -p ../../build_ab/$NAME
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $FLAGS \
  -c sparse_search.cu -o ../../build_ab/$NAME/sparse_search.o 2> ../../build_ab/$NAME/ptxas.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_ab/libsgs_$NAME.so build/hamiltonian.o build/common.o build/capi.o build/nccl_glue.o \
  ../../build_ab/$NAME/sparse_search.o build/group_ops.o build/klocal_search.o -ldl -cudart static
grep -A2 "k_dfsILi[12]ELb0ELb0" ../../build_ab/$NAME/ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
