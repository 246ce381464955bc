#!/bin/bash
# round 2, GPU call d: step_02 kernels (walking regrid v2, batched smoothing)
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py tests/test_cli_gpu.py -m gpu -q --tb=short -x > gpurun_out/tests_r2d.log 2>&1
tail -4 gpurun_out/tests_r2d.log
for v in base rg2; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py > gpurun_out/step02_r2d_$v.json 2> gpurun_out/step02_r2d_$v.err
  echo "== $v $(grep -o '"smoothing": {"ms": [0-9.]*\|"regridding": {"ms": [0-9.]*\|"frac_of_peak": [0-9.]*' gpurun_out/step02_r2d_$v.json | tr '\n' ' ')"
done
unset PGW_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:regrid_walk -c 1 -o gpurun_out/prof_regrid_r2c -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2c.log 2>&1; tail -1 gpurun_out/ncu_regrid_r2c.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:smooth_kernel -c 1 -o gpurun_out/prof_smooth_r2b -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_smooth_r2b.log 2>&1; tail -1 gpurun_out/ncu_smooth_r2b.log | cut -c1-200
