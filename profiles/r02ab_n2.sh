#!/bin/bash
# Round 2: the back-to-back protocol at N = 2 (peer-memory exchange inside one graph of all steps), and the NCCL arm.
OUT=gpurun_out; TAG=r02ab; mkdir -p $OUT; rm -f $OUT/*.ncu-rep
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
  bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_n2_peer.json 2> $OUT/${TAG}_n2_peer.err
echo "rc=$?"; tail -c 300 $OUT/${TAG}_n2_peer.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 \
  bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-secondaries --exchange nccl > $OUT/${TAG}_n2_nccl.json 2> $OUT/${TAG}_n2_nccl.err
echo "rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02ab_n2_peer.json', 'gpurun_out/r02ab_n2_nccl.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, d['value'], d['ms_per_step'], d['device_step'].get('l2'), json.dumps(d['device_step'].get('back_to_back'))[:400], d['device_step'].get('flushed_step_ms_min_med_max'), d['exchange'], d['e2e']['value'])
            print('notes', d['notes'])
            for k, v in d['workloads'].items():
                print(k, v.get('value'), v.get('ms_per_step'), v.get('l2'), json.dumps(v.get('back_to_back'))[:300], v.get('flushed_step_ms_min_med_max'), v.get('error'), v.get('exchange'))
PY
