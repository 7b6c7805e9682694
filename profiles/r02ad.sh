#!/bin/bash
# Round 2: with the candidate-list prefetch on by default: the whole GPU suite, the KD caster with / without the same
# prefetch in its marching search, then the contract command as the driver runs it (N = 1).
OUT=gpurun_out; TAG=r02ad; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -3 $OUT/${TAG}_pytest.txt
for V in base kdpf; do
  LIBV=mcl_3dl_b200/libmcl3dl_b200.so; [ $V != base ] && LIBV=mcl_3dl_b200/libmcl3dl_b200_$V.so
  MCL3DL_LIB=$PWD/$LIBV timeout 300 python bench.py --no-cpu-baseline --no-secondaries --workload c3 --raycaster kd --steps 100 --warmup 5 \
      > $OUT/${TAG}_${V}_c3kd.json 2> $OUT/${TAG}_${V}_c3kd.err
done
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench rc=$?"
python - <<'PY' | tee gpurun_out/r02ad_summary.txt
import json
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
for v in ('base', 'kdpf'):
    d = load('gpurun_out/r02ad_%s_c3kd.json' % v)
    b = d['device_step']['back_to_back']
    print(v, 'c3_kd b2b us %.2f' % (1e3 * d['ms_per_step']), ['%.2f' % (1e3 * x) for x in b['repeats_ms_per_step']], 'flush med %.2f' % (1e3 * d['device_step']['flushed_step_ms_min_med_max'][1]))
d = load('gpurun_out/r02ad_bench_default.json')
print('c2', d['value'], 'us %.2f' % (1e3 * d['ms_per_step']), 'e2e', d['e2e']['value'], 'us %.2f' % (1e3 * d['e2e']['ms_per_step']), 'roofline frac %.3f dram_frac %s' % (d['roofline']['frac'], d['roofline']['dram_frac']))
print('cpu', d['cpu_baseline'])
print('clocks', d['clocks'])
for k, v in d['workloads'].items():
    print(k, v.get('value'), 'us', 1e3 * (v.get('ms_per_step') or 0), v.get('error'), (v.get('roofline') or {}).get('kernel_ms_all'), (v.get('e2e') or {}).get('value'))
PY
