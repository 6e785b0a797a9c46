#!/bin/bash
P=${1:-r02_s7}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -5 $O/${P}_gpu_tests.log
timeout 600 python scripts/quick_c3full.py > $O/${P}_c3full.log 2>&1
cat $O/${P}_c3full.log
timeout 1200 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"
echo done
