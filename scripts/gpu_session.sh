#!/bin/bash
P=${1:-r02_s50}
O=gpurun_out
mkdir -p $O
python scripts/quick_c3full_e2e.py 2>&1 | tee $O/${P}_c3full_e2e_taper.log
timeout 900 python -m pytest tests/test_gpu_streaming.py -m gpu -q -x 2>&1 | tail -2
python scripts/check_streaming_full.py 2>&1 | grep "equal\|call 1" | tee $O/${P}_streaming_full.log
