#!/bin/bash
# round 2, GPU call i: smoothing with two points per thread, step_02 file -> file, launch list of the bench
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x > gpurun_out/tests_r2i.log 2>&1; tail -3 gpurun_out/tests_r2i.log
python tests/bench_step02.py > gpurun_out/step02_r2i.json 2> gpurun_out/step02_r2i.err
echo "== step02 $(grep -o '"smoothing": {"ms": [0-9.]*\|"regridding": {"ms": [0-9.]*\|"frac_of_peak": [0-9.]*' gpurun_out/step02_r2i.json | tr '\n' ' ')"; tail -2 gpurun_out/step02_r2i.err
python tests/bench_step02_files.py > gpurun_out/step02_files_r2.json 2> gpurun_out/step02_files_r2.err; tail -c 1500 gpurun_out/step02_files_r2.json; tail -3 gpurun_out/step02_files_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pgw_|time_mean' -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 0 --streams 1 > gpurun_out/ncu_launches_r2.log 2>&1; grep -c pgw_ gpurun_out/launches_r2.csv
ncu --set full --clock-control none --import-source on -k regex:smooth_kernel -c 1 -o gpurun_out/prof_smooth_r2c -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_smooth_r2c.log 2>&1; tail -1 gpurun_out/ncu_smooth_r2c.log | cut -c1-200
