python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02x_pytest.log 2>&1; tail -2 gpurun_out/r02x_pytest.log
for w in c2 c5; do
  python bench.py --workload $w --no-cpu-baseline --no-secondaries --steps 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('noinline $w', 'ms/step %.4f'%d['ms_per_step'], {k:round(x,4) for k,x in d['roofline']['kernel_ms_all'].items()}, 'e2e %.1f'%(1e3*d['e2e']['ms_per_step']))
"
done
