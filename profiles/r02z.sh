#!/bin/bash
# Round 2, after the wide cells: (1) the rest of the GPU suite, (2) where the c2 launch's 20 us go (cold / code-warm /
# warm), (3) fresh ncu captures of the likelihood kernel that now runs (lik_kernel_nf<..,0,0>) on c2 and c5.
OUT=gpurun_out; TAG=r02z; mkdir -p $OUT
rm -f $OUT/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt
tail -4 $OUT/${TAG}_pytest.txt
timeout 300 python profiles/r02z_cold.py c2 > $OUT/${TAG}_cold.txt 2>&1; tail -2 $OUT/${TAG}_cold.txt
NCU="ncu --clock-control none"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondaries --no-graph"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $OUT/${TAG}_launches_c2.csv $B > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 $B > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c5 $B --workload c5 > /dev/null 2>&1
ls -la $OUT | grep $TAG
