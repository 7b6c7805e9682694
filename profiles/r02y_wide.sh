#!/bin/bash
# Round 2, wide cells of the NN field: parity suite, then c2 / c5 through bench.py (no CPU baseline leg).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02y_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest.txt
tail -5 gpurun_out/r02y_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --steps 200 --warmup 10 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err
tail -c 3000 gpurun_out/r02y_bench.json
