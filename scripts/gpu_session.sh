#!/bin/bash
P=${1:-r02_s5}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -5 $O/${P}_gpu_tests.log
echo done
