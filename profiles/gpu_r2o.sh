#!/bin/bash
# round 2, GPU call o: regrid walk kernel with float64 staging (conversion once per chunk)
cd "$(dirname "$0")/.."
for v in base rf64; do
  unset PGW_B200_LIB
  if [ $v != base ]; then export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=line -x -k "regrid" 2>&1 | tail -1
  python tests/bench_step02.py --reps 7 > gpurun_out/step02_r2o_$v.json 2> gpurun_out/step02_r2o_$v.err
  echo "== $v $(grep -o '"regridding": {"ms": [0-9.]*' gpurun_out/step02_r2o_$v.json)"; tail -1 gpurun_out/step02_r2o_$v.err | cut -c1-200
done
