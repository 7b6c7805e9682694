#!/bin/bash
# Round 2: page-locked caller arrays (mcl3dl_host_alloc): tests, then e2e of c2 / c5 through bench.py.
OUT=gpurun_out; TAG=r02ae; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapter.py tests/test_gpu_resident.py -q -m gpu > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -4 $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
MCL3DL_DIRECT_MIN_KB=-1 timeout 900 python bench.py --no-cpu-baseline --no-secondaries --workload c5 --steps 50 > $OUT/${TAG}_bench_c5_staged.json 2> $OUT/${TAG}_bench_c5_staged.err
python - <<'PY' | tee gpurun_out/r02ae_summary.txt
import json
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
d = load('gpurun_out/r02ae_bench.json')
print('c2 value', d['value'], 'e2e', d['e2e']['value'], 'us %.2f' % (1e3 * d['e2e']['ms_per_step']), d['e2e']['last_call_device_ms'], 'fused us %.2f' % (1e3 * d['e2e']['fused_weight_update']['ms_per_step']))
for k, v in d['workloads'].items():
    e = v.get('e2e') or {}
    print(k, v.get('value'), v.get('error'), 'e2e', e.get('value'), 'us', 1e3 * (e.get('ms_per_step') or 0), e.get('last_call_device_ms'))
d = load('gpurun_out/r02ae_bench_c5_staged.json')
print('c5 staged (MCL3DL_DIRECT_MIN_KB=-1): e2e', d['e2e']['value'], 'us %.2f' % (1e3 * d['e2e']['ms_per_step']))
PY
