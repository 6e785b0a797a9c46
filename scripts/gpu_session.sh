#!/bin/bash
# One GPU-box session: probe, smoke, GPU tests, bench, sweeps.  Logs go to gpurun_out/ (prefix $1).
P=${1:-r02_s1}
O=gpurun_out
mkdir -p $O
{ nproc; free -g; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; lscpu | grep -i -E "numa|model name|socket|thread"; nvidia-smi topo -m; } > $O/${P}_probe.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1 || { echo "smoke failed" >> $O/${P}_smoke.log; }
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
timeout 900 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"
timeout 600 python scripts/quick_e2e.py > $O/${P}_e2e.log 2>&1
QB=8388608 timeout 600 python scripts/ab_variants.py scripts/quick_lv.py 2 > $O/${P}_ab_lv.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${P}_ncu_bench.log 2>&1
echo done
