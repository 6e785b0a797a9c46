#!/bin/bash
P=${1:-r02_s4}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${P}_gpu_tests.log 2>&1
tail -15 $O/${P}_gpu_tests.log
for v in 0 1; do CYRK_EV2=$v timeout 600 python scripts/quick_c3full.py >> $O/${P}_c3full.log 2>&1; done
cat $O/${P}_c3full.log
echo done
