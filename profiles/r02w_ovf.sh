for v in 0 1; do
  if [ $v = 1 ]; then export MCL3DL_NF_IGNORE_OVF=1; else unset MCL3DL_NF_IGNORE_OVF; fi
  for w in c2 c5; do
  python bench.py --workload $w --no-cpu-baseline --no-secondaries --steps 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ignore_ovf $v $w', 'ms/step %.4f'%d['ms_per_step'], {k:round(x,4) for k,x in d['roofline']['kernel_ms_all'].items()}, 'e2e %.1f'%(1e3*d['e2e']['ms_per_step']))
"
  done
done
