#!/bin/bash
# Round 2, first GPU call (1 GPU): close the evidence holes of VERDICT r01 before touching a kernel.
#   gpurun --timeout 1500 -- 'bash profiles/r02a.sh'
TAG=r02a
OUT=gpurun_out
mkdir -p $OUT
# (a) parity suite: f3 un-xfailed, C4/C5-shape oracle checks, fused-update records vs the oracle
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
# (b) every prepared variant byte for byte against the r01x engine + kernel / host-call times
timeout 400 python profiles/ab_variants.py --out $OUT/${TAG}_ab.jsonl --budget 300 --calls 30 > $OUT/${TAG}_ab.log 2>&1
# (c) f2 parity through the one-synchronise fused update
MCL3DL_UPDATE_ONE_SYNC=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused" > $OUT/${TAG}_pytest_one_sync.log 2>&1
tail -2 $OUT/${TAG}_pytest_one_sync.log
# (d) contract bench lines
for w in c2 c3 c5; do
  python bench.py --workload $w --cpu-seconds 4 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
python bench.py --workload c3 --raycaster kd --no-cpu-baseline > $OUT/${TAG}_bench_c3_kd.json 2> $OUT/${TAG}_bench_c3_kd.err
# (e) ncu of the SHIPPED kernels: launch list + full sets (c2 lik, c3 beam DDA, c3 beam KD, c5 lik + beam)
NCU="ncu --clock-control none"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv $B > /dev/null 2>&1
$NCU --set full --import-source on -k regex:lik_kernel_wi -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 $B > /dev/null 2>&1
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_c3 $B --workload c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_kd_c3 $B --workload c3 --raycaster kd > /dev/null 2>&1
$NCU --set full --import-source on -k regex:lik_kernel_wi -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c5 $B --workload c5 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_c5 $B --workload c5 > /dev/null 2>&1
ls -la $OUT | tail -30
