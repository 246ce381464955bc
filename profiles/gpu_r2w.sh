#!/bin/bash
# round 2, GPU call w: compiler-flag variants of the library (ptxas expensive optimisations, fast math)
cd "$(dirname "$0")/.."
B="python bench.py --steps 200 --warmup 5 --no-cpu --e2e-steps 0 --streams 1"
for v in base xo fm base xo fm; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  $B > gpurun_out/var_${v}_w.log 2>&1
  echo "== $v $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/var_${v}_w.log | tr '\n' ' ')"
done
export PGW_B200_LIB=$PWD/scratch/lib_fm.so
python -m pytest tests/test_timestep_gpu.py -m gpu -q --tb=line -x 2>&1 | tail -2
