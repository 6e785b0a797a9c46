#!/bin/bash
# standard single-GPU session: smoke, the GPU test suite, the default bench line
P=${1:-r02_last}
O=gpurun_out
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${P}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
timeout 1500 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"; tail -2 $O/${P}_bench.err
