#!/bin/bash
P=${1:-r02_s34}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_streaming.py tests/test_gpu_boundary.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -4 $O/${P}_gpu_tests.log
python scripts/quick_c3full_e2e.py 2>&1 | head -1 | tee $O/${P}_c3full_e2e.log
python scripts/check_streaming_full.py 2>&1 | tail -12 | tee $O/${P}_streaming_full.log
echo done
