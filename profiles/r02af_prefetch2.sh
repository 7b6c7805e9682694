#!/bin/bash
# Round 2: more prefetch variants of lik_kernel_nf: next trip's poses (pp), + candidate depth 5 (pp5), depth 1 (d1).
OUT=gpurun_out; TAG=r02af; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
for V in base pp pp5 d1; do
  LIBV=mcl_3dl_b200/libmcl3dl_b200.so; [ $V != base ] && LIBV=mcl_3dl_b200/libmcl3dl_b200_$V.so
  for W in c2 c5; do
    MCL3DL_LIB=$PWD/$LIBV timeout 300 python bench.py --no-cpu-baseline --no-secondaries --workload $W --steps 100 --warmup 5 \
      > $OUT/${TAG}_${V}_${W}.json 2> $OUT/${TAG}_${V}_${W}.err
  done
done
python - <<'PY' | tee gpurun_out/r02af_summary.txt
import json
for v in ('base', 'pp', 'pp5', 'd1'):
    for w in ('c2', 'c5'):
        try:
            for l in open('gpurun_out/r02af_%s_%s.json' % (v, w)):
                if l.startswith('{'):
                    d = json.loads(l)
                    b = d['device_step']['back_to_back']
                    print(v, w, 'b2b us %.2f' % (1e3 * d['ms_per_step']), 'repeats', ['%.2f' % (1e3 * x) for x in b['repeats_ms_per_step']],
                          'flush med %.2f' % (1e3 * d['device_step']['flushed_step_ms_min_med_max'][1]),
                          'kern', {k: round(1e3 * x, 2) for k, x in d['roofline']['kernel_ms_all'].items()}, 'e2e us %.2f' % (1e3 * d['e2e']['ms_per_step']))
        except Exception as e:
            print(v, w, 'error', e)
PY
