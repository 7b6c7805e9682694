for sh in 4 3 2 1; do
 for ov in 1 0; do
  MCL3DL_LIK_SHARE=$sh MCL3DL_OVERLAP=$ov python bench.py --workload c5 --no-cpu-baseline --no-secondaries --steps 40 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('share $sh overlap $ov', 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_all'], 'e2e %.1f'%(1e3*d['e2e']['ms_per_step']))
"
 done
done
