#!/bin/bash
# round 2, GPU call j: tuning variants of the step_02 kernels
cd "$(dirname "$0")/.."
for v in base sb16 sb4 st256 sc16 sc32 rc32 rc8; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py --reps 7 > gpurun_out/step02_r2j_$v.json 2> gpurun_out/step02_r2j_$v.err
  echo "== $v $(grep -o '"smoothing": {"ms": [0-9.]*\|"regridding": {"ms": [0-9.]*' gpurun_out/step02_r2j_$v.json | tr '\n' ' ')"; tail -1 gpurun_out/step02_r2j_$v.err | cut -c1-200
done
