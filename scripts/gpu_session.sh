#!/bin/bash
P=${1:-r02_s53}
O=gpurun_out
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${P}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
timeout 1500 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"; tail -2 $O/${P}_bench.err
QN=10000 QB=592 python scripts/quick_c4.py > $O/${P}_plain_c4ref.log 2>&1 &&
QN=10000 QB=592 ncu --set full --clock-control none --import-source on -k regex:rk_cta2_kernel -s 1 -c 1 -f -o $O/${P}_prof_c4ref python scripts/quick_c4.py > $O/${P}_ncu_c4ref.log 2>&1
echo "c4ref rc=$?"
python scripts/ncu_kernel_summary.py $O/${P}_prof_c4ref.ncu-rep > $O/${P}_ncu_c4ref_summary.txt 2>&1; head -40 $O/${P}_ncu_c4ref_summary.txt
QB=1184 ncu --set full --clock-control none --import-source on -k regex:rk_cta2_kernel -s 1 -c 1 -f -o $O/${P}_prof_c4 python scripts/quick_c4.py > $O/${P}_ncu_c4.log 2>&1
python scripts/ncu_kernel_summary.py $O/${P}_prof_c4.ncu-rep > $O/${P}_ncu_c4_summary.txt 2>&1; head -24 $O/${P}_ncu_c4_summary.txt
