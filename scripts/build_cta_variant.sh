#!/bin/bash
# usage: scripts/build_cta_variant.sh NAME "EXTRA NVCC FLAGS": variants/lib_NAME.so = the current library with
# inst_cta0_fast.o (largey, production build) recompiled with the extra flags (A/B runs of the CTA-per-trajectory kernels: scripts/ab_variants.py)
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
B=cyrk_b200/csrc/build_$NAME
mkdir -p $B variants
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --compress-mode=size -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     -fmad=true -DCYRK_FP_NS=fast -DCYRK_CTA_PART=0 $EXTRA -c cyrk_b200/csrc/inst_cta.cu -o $B/inst_cta0_fast.o
OBJS=$(ls cyrk_b200/csrc/build/*.o | grep -v inst_cta0_fast.o)
nvcc -gencode arch=compute_100a,code=sm_100a --compress-mode=size -shared -cudart static -o variants/lib_$NAME.so $OBJS $B/inst_cta0_fast.o -ldl
cuobjdump -res-usage $B/inst_cta0_fast.o 2>/dev/null | grep -A1 "cta2" | grep -B1 "REG" | paste - - | sed 's/Fatbin.*//' | awk '{print $2, $3,$4,$5,$6}' | c++filt | grep "ProbLargeY, 4, 8, false"
