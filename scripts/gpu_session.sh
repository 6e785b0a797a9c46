#!/bin/bash
P=${1:-r02_s22}
O=gpurun_out
mkdir -p $O
echo "== new psin (psin_pi + check)" > $O/${P}_psin_ab.log
python scripts/quick_c3full.py >> $O/${P}_psin_ab.log 2>&1
python scripts/quick_c3.py 2>&1 | head -3 >> $O/${P}_psin_ab.log
echo "== old psin (two half-range polynomials)" >> $O/${P}_psin_ab.log
CYRK_B200_LIB=$PWD/variants/lib_oldsin.so CYRK_B200_NVRTC_DEFINES="-DCYRK_PSIN_PI=0" python scripts/quick_c3full.py >> $O/${P}_psin_ab.log 2>&1
CYRK_B200_LIB=$PWD/variants/lib_oldsin.so CYRK_B200_NVRTC_DEFINES="-DCYRK_PSIN_PI=0" python scripts/quick_c3.py 2>&1 | head -3 >> $O/${P}_psin_ab.log
cat $O/${P}_psin_ab.log
python scripts/quick_c4_methods.py > $O/${P}_c4_methods.log 2>&1; tail -12 $O/${P}_c4_methods.log
timeout 1800 python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py tests/test_gpu_events_two_pass.py -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -5 $O/${P}_gpu_tests.log
echo done
