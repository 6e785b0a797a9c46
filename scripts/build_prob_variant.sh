#!/bin/bash
# usage: scripts/build_prob_variant.sh NAME PROB RHS_ID "EXTRA NVCC FLAGS": variants/lib_NAME.so = the current library with
# inst_PROB_fast.o recompiled with the extra flags (A/B runs of one built-in problem's thread-per-trajectory kernels)
set -e
cd "$(dirname "$0")/.."
NAME=$1; PROB=$2; ID=$3; EXTRA=$4
B=cyrk_b200/csrc/build_$NAME
mkdir -p $B variants
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --compress-mode=size -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     -fmad=true -DCYRK_INST_PROB=$PROB -DCYRK_INST_ID=$ID -DCYRK_FP_NS=fast -DCYRK_FP_MODE=FP_FAST $EXTRA -c cyrk_b200/csrc/inst_problem.cu -o $B/inst_${PROB}_fast.o
OBJS=$(ls cyrk_b200/csrc/build/*.o | grep -v inst_${PROB}_fast.o)
nvcc -gencode arch=compute_100a,code=sm_100a --compress-mode=size -shared -cudart static -o variants/lib_$NAME.so $OBJS $B/inst_${PROB}_fast.o -ldl
