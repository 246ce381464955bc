#!/bin/bash
# round 2, GPU call c: whole GPU suite with the packed-f32x2 kernel, unroll/slot/stream variants, new regrid kernel
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests_r2c.log 2>&1
tail -8 gpurun_out/tests_r2c.log
B="python bench.py --steps 100 --warmup 5 --no-cpu --e2e-steps 0"
run() { # tag streams
  $B --streams $2 > gpurun_out/var_$1_s$2.log 2>&1
  echo "== $1 streams=$2 $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/var_$1_s$2.log | tr '\n' ' ')"
}
unset PGW_B200_LIB
for s in 1 2 3; do run x2 $s; done
for v in u2s u2p u2 sl5; do export PGW_B200_LIB=$PWD/scratch/lib_$v.so; run $v 2; done
unset PGW_B200_LIB
python tests/bench_step02.py > gpurun_out/step02_r2c.json 2> gpurun_out/step02_r2c.err; cut -c300-1100 gpurun_out/step02_r2c.json; tail -3 gpurun_out/step02_r2c.err
PGW_REGRID_PATH=rows python tests/bench_step02.py 2>/dev/null | grep -o '"regridding": {"ms": [0-9.]*'
ncu --set full --clock-control none --import-source on -k regex:regrid_walk -c 1 -o gpurun_out/prof_regrid_r2b -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2b.log 2>&1; tail -1 gpurun_out/ncu_regrid_r2b.log | cut -c1-200
