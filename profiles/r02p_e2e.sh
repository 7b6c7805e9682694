OUT=gpurun_out
python -m pytest tests -x -q -m gpu > $OUT/r02p_pytest.log 2>&1; tail -3 $OUT/r02p_pytest.log
for v in 262144 0; do
  MCL3DL_STAGE_IN_MAX=$v python bench.py --no-cpu-baseline --no-secondaries --steps 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('stage_in_max $v', 'kern %.4f'%d['roofline']['kernel_ms'], 'e2e %.1f us'%(1e3*d['e2e']['ms_per_step']), d['e2e']['last_call_device_ms'], 'fused %.1f us'%(1e3*d['e2e']['fused_weight_update']['ms_per_step']))
"
  MCL3DL_STAGE_IN_MAX=$v python bench.py --no-cpu-baseline --no-secondaries --steps 50 --workload c1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('c1 stage_in_max $v', 'e2e %.1f us'%(1e3*d['e2e']['ms_per_step']), 'fused %.1f us'%(1e3*d['e2e']['fused_weight_update']['ms_per_step']))
"
done
