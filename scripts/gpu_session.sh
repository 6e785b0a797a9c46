#!/bin/bash
P=${1:-r02_s31}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_medium_ny.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -8 $O/${P}_gpu_tests.log
timeout 900 python scripts/bench_medium_ny.py 2>&1 | tail -5 | tee $O/${P}_medium_ny.log
echo done
