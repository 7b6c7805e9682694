#!/bin/bash
# Round 2: the back-to-back (rotating inputs) protocol next to the flush protocol, N = 1.
OUT=gpurun_out; TAG=r02aa; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
timeout 900 python bench.py --no-cpu-baseline --steps 200 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02aa_bench.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('c2', d['value'], d['ms_per_step'], d['device_step'].get('l2'), json.dumps(d['device_step'].get('back_to_back')), d['device_step'].get('flushed_step_ms_min_med_max'))
        print('notes', d['notes'])
        for k, v in d['workloads'].items():
            print(k, v.get('value'), v.get('ms_per_step'), v.get('l2'), json.dumps(v.get('back_to_back')), v.get('flushed_step_ms_min_med_max'), v.get('error'))
PY
