#!/bin/bash
# final captures of the round: every profiled command first runs plain in the same call
P=${1:-r02_final}
O=gpurun_out
mkdir -p $O
python scripts/prof_c5.py > $O/${P}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk_thread_kernel -s 1 -c 1 -f -o $O/${P}_prof_c5 python scripts/prof_c5.py > $O/${P}_ncu_c5.log 2>&1
echo "c5 rc=$?"
python scripts/ncu_kernel_summary.py $O/${P}_prof_c5.ncu-rep > $O/${P}_ncu_c5_summary.txt 2>&1; head -24 $O/${P}_ncu_c5_summary.txt
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${P}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_ncu_launches_C5.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${P}_ncu_bench.log 2>&1
echo "launches rc=$?"
QB=1048576 python scripts/prof_c3full.py > $O/${P}_plain_c3.log 2>&1 &&
QB=1048576 ncu --set full --clock-control none --import-source on -k regex:"cyrk_user_kernel|event_roots" -c 2 -f -o $O/${P}_prof_c3full python scripts/prof_c3full.py > $O/${P}_ncu_c3.log 2>&1
echo "c3 rc=$?"
python scripts/ncu_kernel_summary.py $O/${P}_prof_c3full.ncu-rep cyrk_user_kernel > $O/${P}_ncu_c3full_evdetect.txt 2>&1; head -24 $O/${P}_ncu_c3full_evdetect.txt
python scripts/prof_c2tp.py 2 > $O/${P}_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"step_dense|rk_thread_kernel" -s 2 -c 2 -f -o $O/${P}_prof_c2full python scripts/prof_c2tp.py 2 > $O/${P}_ncu_c2.log 2>&1
echo "c2 rc=$?"
python scripts/ncu_kernel_summary.py $O/${P}_prof_c2full.ncu-rep step_dense > $O/${P}_ncu_c2_step_dense.txt 2>&1
python scripts/ncu_kernel_summary.py $O/${P}_prof_c2full.ncu-rep rk_thread_kernel > $O/${P}_ncu_c2_steps_solve.txt 2>&1
ls -la $O/${P}*.ncu-rep
