#!/bin/bash
# 8-GPU box: the exchange test at world 8, then the contract bench at N = 8 and N = 4 (and the NCCL comparison at 8)
OUT=gpurun_out; TAG=r02n; mkdir -p $OUT
python -m pytest tests/test_gpu_exchange.py -x -q -m gpu > $OUT/${TAG}_pytest_exchange.log 2>&1; tail -3 $OUT/${TAG}_pytest_exchange.log
PORT=29800
for N in 8 4; do
  PORT=$((PORT+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_n${N}_peer.json 2> $OUT/${TAG}_n${N}_peer.err
  tail -c 200 $OUT/${TAG}_n${N}_peer.err
done
PORT=$((PORT+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --exchange nccl > $OUT/${TAG}_n8_nccl.json 2> $OUT/${TAG}_n8_nccl.err
ls -la $OUT | tail -8
