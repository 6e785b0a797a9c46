#!/bin/bash
P=${1:-r02_s28}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests/test_gpu_largey.py tests/test_user_rhs.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -8 $O/${P}_gpu_tests.log
python scripts/quick_c4.py 2>&1 | tee $O/${P}_c4.log
echo done
