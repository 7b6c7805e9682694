#!/bin/bash
# final scaling sweep of the round on one 8-GPU box: contract command at N = 1, 2, 4, 8 (+ reference arm at N = 1)
OUT=gpurun_out; TAG=r02t; mkdir -p $OUT
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; tail -c 200 $OUT/${TAG}_n1.err
PORT=29900
for N in 2 4 8; do
  PORT=$((PORT+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_n${N}.json 2> $OUT/${TAG}_n${N}.err
  tail -c 200 $OUT/${TAG}_n${N}.err
done
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_n1_reference.json 2>/dev/null
ls -la $OUT | grep $TAG
