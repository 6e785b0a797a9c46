#!/bin/bash
P=${1:-r02_s21}
O=gpurun_out
mkdir -p $O



timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -15 $O/${P}_gpu_tests.log
echo done
