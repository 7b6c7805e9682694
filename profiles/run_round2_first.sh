#!/bin/bash
# The first GPU calls of round 2, in order.  Everything marked (prepared) was written after round 1's GPU budget ran
# out: host-verified where the code is per-thread, otherwise never executed.  TAG names the output files.
#
#   HERE, before the first call (nvcc, no GPU):   python profiles/ab_variants.py --build-variants
#
#   1 GPU:   gpurun --timeout 900 -- 'bash profiles/run_round2_first.sh one r02a'
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_round2_first.sh two r02a'
MODE=${1:-one}
TAG=${2:-r02a}
OUT=gpurun_out
mkdir -p $OUT
if [ "$MODE" = one ]; then
  # (a) the parity suite on today's defaults (near-field screens, zero-copy records, timing events off)
  python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
  # (b) every prepared variant, byte for byte against the r01x engine, with kernel / host-call / fused-update times
  python profiles/ab_variants.py --out $OUT/${TAG}_ab.jsonl --budget 240 --calls 30
  # (c) the f2 parity tests through the one-synchronise fused update (prepared)
  MCL3DL_UPDATE_ONE_SYNC=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or update" > $OUT/${TAG}_pytest_one_sync.log 2>&1
  tail -2 $OUT/${TAG}_pytest_one_sync.log
  # (d) contract bench lines of the headline workloads
  for w in c2 c3 c5 c1; do
    python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
  done
  python bench.py --workload c3 --raycaster kd --no-cpu-baseline > $OUT/${TAG}_bench_c3_kd.json 2> $OUT/${TAG}_bench_c3_kd.err
  # (e) ncu of the likelihood kernel with the screens (the committed capture predates them): launch list + full set
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:lik_kernel_wi -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
else
  # 2 GPUs: NCCL all-gather (contract) vs CUDA-graph replay (prepared) vs the peer-memory exchange kernel (prepared)
  PORT=29700
  for w in c2 c5; do
    for v in "" "--graph" "--exchange peer"; do
      PORT=$((PORT+1))
      name=$(echo "nccl$v" | tr -d ' -')
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus 2 --workload $w --steps 30 --warmup 5 --no-cpu-baseline $v \
        > $OUT/${TAG}_${w}_n2_$name.json 2> $OUT/${TAG}_${w}_n2_$name.err
      tail -c 300 $OUT/${TAG}_${w}_n2_$name.err
    done
  done
  python profiles/bench_summary.py $OUT/${TAG}_*_n2_*.json
fi
