#!/bin/bash
# round 2, GPU call n: smoothing with the harmonic tables in constant memory, resident-warp variants
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "smooth or regrid" > gpurun_out/tests_r2n.log 2>&1; tail -2 gpurun_out/tests_r2n.log
for v in base shared sm10 sm12 sm16; do
  unset PGW_B200_LIB PGW_SMOOTH_TABLE
  if [ $v = shared ]; then export PGW_SMOOTH_TABLE=shared; elif [ $v != base ]; then export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py --reps 7 > gpurun_out/step02_r2n_$v.json 2> gpurun_out/step02_r2n_$v.err
  echo "== $v $(grep -o '"smoothing": {"ms": [0-9.]*' gpurun_out/step02_r2n_$v.json)  $(grep -o '"max_abs_err_vs_oracle": [0-9.e-]*' gpurun_out/step02_r2n_$v.json | sed -n 2p)"; tail -1 gpurun_out/step02_r2n_$v.err | cut -c1-200
done
