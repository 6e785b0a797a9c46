#!/bin/bash
P=${1:-r02_s3}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
timeout 900 python scripts/bench_medium_ny.py > $O/${P}_medium_ny.log 2>&1
tail -13 $O/${P}_medium_ny.log
echo done
