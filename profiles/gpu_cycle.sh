#!/bin/bash
# One GPU round trip: parity tests, a short bench, and (optionally) an ncu capture of the column kernel.
#   gpurun --timeout 900 -- 'bash profiles/gpu_cycle.sh v11'
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/plain_$tag.log 2>&1
grep -o "\"kernel_ms\": [0-9.]*\|\"ms_per_step\": [0-9.]*" gpurun_out/plain_$tag.log
if [ "$2" != "noprof" ]; then
ncu --set full --clock-control none --import-source on -k regex:pgw_column -s 3 -c 1 \
    -o gpurun_out/prof_column_$tag -f python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log | cut -c1-120
fi
