#!/bin/bash
# Run ON THE GPU BOX (under gpurun): bench lines for every single-GPU workload + ncu evidence.
# Usage: gpurun --timeout 1500 -- 'bash profiles/run_profiles.sh r01'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
for w in c2 c3 c1 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
python bench.py --workload c2 --spread --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c2_spread.json 2> $OUT/${TAG}_bench_c2_spread.err
python bench.py --workload c3 --spread --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c3_spread.json 2> $OUT/${TAG}_bench_c3_spread.err
python bench.py --impl reference --workload c2 --steps 3 --warmup 1 > $OUT/${TAG}_bench_c2_reference.json 2> $OUT/${TAG}_bench_c2_reference.err
# every launch with its device time (cold-cache, serialised): compare shares
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_c3.log 2>&1
# the two hot kernels, full sections, once
ncu --set full --clock-control none --import-source on -k regex:lik_kernel -s 4 -c 2 -f -o $OUT/${TAG}_prof_lik_c2 \
    python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 4 -c 2 -f -o $OUT/${TAG}_prof_beam_c3 \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lik_kernel -s 4 -c 2 -f -o $OUT/${TAG}_prof_lik_c2_spread \
    python bench.py --workload c2 --spread --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full_c2s.log 2>&1
ls -la $OUT | tail -30
for w in c2 c3 c1 c5 c2_spread c3_spread c2_reference; do echo "== $w"; head -c 600 $OUT/${TAG}_bench_$w.json; echo; tail -2 $OUT/${TAG}_bench_$w.err; done
