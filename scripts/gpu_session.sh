#!/bin/bash
P=${1:-r02_s49}
O=gpurun_out
mkdir -p $O
for n in 10000 8192 6000 12288; do QN=$n QB=2048 python scripts/ab_variants.py scripts/quick_c4.py 2>&1 | grep -v "largey_simple\|RK23" ; done | tee $O/${P}_large_unroll.log
