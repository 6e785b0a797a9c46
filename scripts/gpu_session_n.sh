#!/bin/bash
# multi-GPU bench session: $1 = N, $2 = prefix
N=${1:-2}
P=${2:-r02_n$N}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err
echo "bench rc=$?"; tail -c 600 $O/${P}_bench.json
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/${P}_bench_ref.json 2> $O/${P}_bench_ref.err
echo "ref rc=$?"; nproc; free -g | head -2
if [ "$N" -ge 4 ]; then
  # BASELINE configs 2 and 3 at their stated size (4 Mi trajectories) sharded over the box
  PER=$((4194304 / N))
  timeout 1200 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload C2full --batch $PER --no-extras > $O/${P}_bench_C2full.json 2> $O/${P}_bench_C2full.err
  echo "C2full rc=$?"; tail -c 400 $O/${P}_bench_C2full.err
  timeout 1200 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --workload C3full --batch $PER --no-extras > $O/${P}_bench_C3full.json 2> $O/${P}_bench_C3full.err
  echo "C3full rc=$?"; tail -c 400 $O/${P}_bench_C3full.err
fi
