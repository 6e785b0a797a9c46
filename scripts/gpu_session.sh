#!/bin/bash
P=${1:-r02_s43}
O=gpurun_out
mkdir -p $O
echo "== psin2" > $O/${P}_psin2_ab.log
python scripts/quick_c3full.py 2>&1 | tail -1 >> $O/${P}_psin2_ab.log
python scripts/quick_c3.py 2>&1 | head -6 >> $O/${P}_psin2_ab.log
echo "== two psin calls (CYRK_PSIN2=0; NVRTC kernels only: the detection solve)" >> $O/${P}_psin2_ab.log
CYRK_B200_NVRTC_DEFINES="-DCYRK_PSIN2=0" python scripts/quick_c3full.py 2>&1 | tail -1 >> $O/${P}_psin2_ab.log
cat $O/${P}_psin2_ab.log
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_events_two_pass.py tests/test_user_rhs.py tests/test_gpu_boundary.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
