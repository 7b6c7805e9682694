#!/bin/bash
# fresh ncu --set full captures of the SHIPPED kernels + the launch list of the contract command (1 GPU).
# gpurun brings back at most 64 MiB and a full report with sources is ~12 MB: run it in two parts (a | b).
PART=${1:-a}
OUT=gpurun_out; TAG=r02o; mkdir -p $OUT
NCU="ncu --clock-control none"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondaries --no-graph"
if [ "$PART" = a ]; then
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $OUT/${TAG}_launches_c2.csv $B > /dev/null 2>&1
$NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c2 $B > /dev/null 2>&1
$NCU --set full --import-source on -k regex:lik_kernel_nf -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_c5 $B --workload c5 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_c5 $B --workload c5 > /dev/null 2>&1
else
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_c3 $B --workload c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:beam_kernel_pl -c 1 -s 5 -o $OUT/${TAG}_ncu_beam_kd_c3 $B --workload c3 --raycaster kd > /dev/null 2>&1
MCL3DL_LIK_MODE=field $NCU --set full --import-source on -k regex:lik_kernel_field -c 1 -s 5 -o $OUT/${TAG}_ncu_lik_field_c5 $B --workload c5 > /dev/null 2>&1
fi
ls -la $OUT | grep $TAG
