#!/bin/bash
# Run ON an 8-GPU box: gpurun --gpus 8 --timeout 1200 -- 'bash profiles/run_scaling.sh r01s'
TAG=${1:-r01s}
OUT=gpurun_out
mkdir -p $OUT
PORT=29600
for w in c5 c2; do
  for n in 1 2 4 8; do
    PORT=$((PORT+1))
    if [ $n -eq 1 ]; then
      python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_${w}_n$n.json 2> $OUT/${TAG}_${w}_n$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_${w}_n$n.json 2> $OUT/${TAG}_${w}_n$n.err
    fi
  done
done
python profiles/bench_summary.py $OUT/${TAG}_*.json
