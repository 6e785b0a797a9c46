#!/bin/bash
P=${1:-r02_s8}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${P}_gpu_tests.log 2>&1
tail -5 $O/${P}_gpu_tests.log
timeout 300 python scripts/sanitize_r02.py > $O/${P}_sanitize_plain.log 2>&1 &&
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_r02.py > $O/${P}_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 $O/${P}_sanitize_memcheck.log; tail -3 $O/${P}_sanitize_plain.log
timeout 1200 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"
echo done
