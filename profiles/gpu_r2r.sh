#!/bin/bash
# round 2, GPU call r: smoothing with a cp.async ring for the read pass
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py tests/test_cli_gpu.py -m gpu -q --tb=short -x -k "smooth or step_02" 2>&1 | tail -2
for v in base r16 r64 r32b4; do
  unset PGW_B200_LIB
  if [ $v != base ]; then export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py --reps 7 > gpurun_out/step02_r2r_$v.json 2> gpurun_out/step02_r2r_$v.err
  echo "== $v $(grep -o '"smoothing": {"ms": [0-9.]*' gpurun_out/step02_r2r_$v.json) $(grep -o '"max_abs_err_vs_oracle": [0-9.e-]*' gpurun_out/step02_r2r_$v.json | sed -n 2p)"; tail -1 gpurun_out/step02_r2r_$v.err | cut -c1-200
done
