#!/bin/bash
# round 2, GPU call q: where the smoothing kernel spends its time (read pass only / write pass only / 512-thread CTAs)
cd "$(dirname "$0")/.."
for v in base so1 so2 st512; do
  unset PGW_B200_LIB
  if [ $v != base ]; then export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py --reps 7 > gpurun_out/step02_r2q_$v.json 2> gpurun_out/step02_r2q_$v.err
  echo "== $v $(grep -o '"smoothing": {"ms": [0-9.]*' gpurun_out/step02_r2q_$v.json)"; tail -1 gpurun_out/step02_r2q_$v.err | cut -c1-200
done
