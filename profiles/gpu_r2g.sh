#!/bin/bash
# round 2, GPU call g: regrid walk v4 (pipelined passes, float64 staging), 3 vs 2 CTAs/SM
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x > gpurun_out/tests_r2g.log 2>&1; tail -3 gpurun_out/tests_r2g.log
for v in base rg2; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  python tests/bench_step02.py > gpurun_out/step02_r2g_$v.json 2> gpurun_out/step02_r2g_$v.err
  echo "== $v $(grep -o '"regridding": {"ms": [0-9.]*' gpurun_out/step02_r2g_$v.json)"; tail -2 gpurun_out/step02_r2g_$v.err
done
unset PGW_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:regrid_walk -c 1 -o gpurun_out/prof_regrid_r2e -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2e.log 2>&1; tail -1 gpurun_out/ncu_regrid_r2e.log | cut -c1-200
