#!/bin/bash
# short 8-GPU session: headline (weak + strong + ensemble) and C3full at 4 Mi total
N=${1:-8}
P=${2:-r02_n8b}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"; tail -c 300 $O/${P}_bench.err
PER=$((4194304 / N))
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --workload C3full --batch $PER --no-extras --no-cpu-baseline > $O/${P}_bench_C3full.json 2> $O/${P}_bench_C3full.err
echo "C3full rc=$?"; tail -c 300 $O/${P}_bench_C3full.err
timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --workload C4 --no-extras --no-cpu-baseline > $O/${P}_bench_C4.json 2> $O/${P}_bench_C4.err
echo "C4 rc=$?"; tail -c 300 $O/${P}_bench_C4.err
