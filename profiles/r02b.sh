#!/bin/bash
# Round 2, second GPU call: NN field + folded record exchange.
#   1 GPU :  gpurun --timeout 1200 -- 'bash profiles/r02b.sh one'
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash profiles/r02b.sh two'
MODE=${1:-one}
TAG=${2:-r02c}
OUT=gpurun_out
mkdir -p $OUT
if [ "$MODE" = one ]; then
  python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
  timeout 500 python profiles/ab_variants.py --out $OUT/${TAG}_ab.jsonl --budget 400 --calls 30 > $OUT/${TAG}_ab.log 2>&1
  python bench.py --cpu-seconds 4 > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench_c2.err; tail -c 400 $OUT/${TAG}_bench_c2.err
  NCU="ncu --clock-control none"
  B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondaries --no-graph"
  $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 $B > /dev/null 2>&1
  $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c5 $B --workload c5 > /dev/null 2>&1
  $NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_kd_c3 $B --workload c3 --raycaster kd > /dev/null 2>&1
else
  N=${3:-2}
  python -m pytest tests/test_gpu_exchange.py tests/test_gpu_parity.py -x -q -m gpu -k "exchange or multi_device" > $OUT/${TAG}_pytest_n$N.log 2>&1; tail -3 $OUT/${TAG}_pytest_n$N.log
  PORT=29700
  for v in "" "--exchange nccl" "--no-graph" "--exchange nccl --no-graph"; do
    PORT=$((PORT+1))
    name=$(echo "peer$v" | tr -d ' -')
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline $v \
      > $OUT/${TAG}_n${N}_$name.json 2> $OUT/${TAG}_n${N}_$name.err
    tail -c 300 $OUT/${TAG}_n${N}_$name.err
  done
fi
ls -la $OUT | tail -12
