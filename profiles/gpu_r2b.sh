#!/bin/bash
# round 2, GPU call b: updated parity tests, kernel variants (packed f32x2, streams), step_02 baseline + ncu
cd "$(dirname "$0")/.."
export PGW_REFDTYPES_OUT=$PWD/gpurun_out/refdtypes_lines.jsonl; rm -f $PGW_REFDTYPES_OUT
python -m pytest tests/test_ref_dtypes_gpu.py tests/test_cli_gpu.py tests/test_timestep_gpu.py -m gpu -q --tb=short > gpurun_out/tests_r2b.log 2>&1
tail -15 gpurun_out/tests_r2b.log
B="python bench.py --steps 100 --warmup 5 --no-cpu --e2e-steps 0"
for v in base x2; do
  if [ $v = base ]; then unset PGW_B200_LIB; else export PGW_B200_LIB=$PWD/scratch/lib_$v.so; fi
  for s in 1 2; do
    $B --streams $s > gpurun_out/var_${v}_s$s.log 2>&1
    echo "== $v streams=$s $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"host_us_per_submit": [0-9.]*\|"host_us_fill_args": [0-9.]*' gpurun_out/var_${v}_s$s.log | tr '\n' ' ')"
  done
done
unset PGW_B200_LIB
python tests/bench_step02.py > gpurun_out/step02_r2b.json 2> gpurun_out/step02_r2b.err; cut -c1-900 gpurun_out/step02_r2b.json
ncu --set full --clock-control none --import-source on -k regex:regrid_rows -c 1 -o gpurun_out/prof_regrid_r2a -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2a.log 2>&1; tail -2 gpurun_out/ncu_regrid_r2a.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:smooth_kernel -c 1 -o gpurun_out/prof_smooth_r2a -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_smooth_r2a.log 2>&1; tail -2 gpurun_out/ncu_smooth_r2a.log | cut -c1-200
