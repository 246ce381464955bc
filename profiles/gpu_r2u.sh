#!/bin/bash
# round 2, GPU call u: uniform-level shortcut of the walker (PGW_UNIFORM_TOP) against the default build
cd "$(dirname "$0")/.."
export PGW_B200_LIB=$PWD/scratch/lib_ut.so
python -m pytest tests/test_timestep_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -2
unset PGW_B200_LIB
B="python bench.py --steps 200 --warmup 5 --no-cpu --e2e-steps 0"
for rep in 1 2; do
for v in base ut; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  for s in 1 2; do
    $B --streams $s > gpurun_out/var_${v}_s${s}_u.log 2>&1
    echo "== $v streams=$s $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/var_${v}_s${s}_u.log | tr '\n' ' ')"
  done
done
done
