#!/bin/bash
P=${1:-r02_s14}
O=gpurun_out
mkdir -p $O
python scripts/quick_c4.py > $O/${P}_c4.log 2>&1; cat $O/${P}_c4.log
timeout 1800 python -m pytest tests/test_gpu_largey.py tests/test_user_rhs.py tests/test_gpu_math.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -5 $O/${P}_gpu_tests.log
echo done
