for sh in 4 3 2; do for bm in dq pl; do
  MCL3DL_LIK_SHARE=$sh MCL3DL_BEAM=$bm python bench.py --workload c5e --no-cpu-baseline --no-secondaries --steps 50 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('share $sh beam $bm', 'ms/step %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms_all'].items()}, 'e2e %.1f'%(1e3*d['e2e']['ms_per_step']))
"
done; done
