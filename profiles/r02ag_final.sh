#!/bin/bash
# Round 2, final state: the whole GPU suite, the contract command (N = 1, CPU baseline leg included), the ncu launch list
# of the same command and fresh full captures of the likelihood kernel (c2, c5) with the candidate-list prefetch.
OUT=gpurun_out; TAG=r02ag; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt; tail -3 $OUT/${TAG}_pytest.txt
timeout 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondaries --no-graph"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $OUT/${TAG}_launches_c2.csv $B > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 $B > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c5 $B --workload c5 > /dev/null 2>&1
python - <<'PY' | tee gpurun_out/r02ag_summary.txt
import json
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
d = load('gpurun_out/r02ag_bench_n1.json')
print('c2', d['value'], 'us %.2f' % (1e3 * d['ms_per_step']), 'e2e', d['e2e']['value'], 'us %.2f' % (1e3 * d['e2e']['ms_per_step']), 'fused us %.2f' % (1e3 * d['e2e']['fused_weight_update']['ms_per_step']))
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'dram_frac', 'kernel_ms', 'traffic')})
print('b2b', d['device_step']['back_to_back']['repeats_ms_per_step'], d['device_step']['back_to_back']['same_station_every_step_ms_per_step'], d['device_step']['flushed_step_ms_min_med_max'])
print('cpu', d['cpu_baseline'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'], 'notes', d['notes'])
for k, v in d['workloads'].items():
    e = v.get('e2e') or {}
    print(k, v.get('value'), 'us', 1e3 * (v.get('ms_per_step') or v.get('ms_per_cycle') or 0), v.get('error'), (v.get('roofline') or {}).get('kernel_ms_all'), 'e2e', e.get('value'), 'us', 1e3 * (e.get('ms_per_step') or 0), v.get('ms_per_call'))
PY
ls -la $OUT | grep $TAG
