#!/bin/bash
# One GPU-box session: smoke, GPU tests, bench, checks.  Logs go to gpurun_out/ (prefix $1).
P=${1:-r02_s2}
O=gpurun_out
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1 || { echo "smoke failed" >> $O/${P}_smoke.log; }
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
timeout 900 python scripts/check_streaming_full.py > $O/${P}_streaming.log 2>&1
tail -4 $O/${P}_streaming.log
timeout 1200 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"
echo done
