#!/bin/bash
# round 2, GPU call f: regrid walk v3, column kernel ncu capture + launch list, default bench
cd "$(dirname "$0")/.."
python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x > gpurun_out/tests_r2f.log 2>&1; tail -3 gpurun_out/tests_r2f.log
python tests/bench_step02.py > gpurun_out/step02_r2f.json 2> gpurun_out/step02_r2f.err
echo "== step02 $(grep -o '"smoothing": {"ms": [0-9.]*\|"regridding": {"ms": [0-9.]*\|"frac_of_peak": [0-9.]*' gpurun_out/step02_r2f.json | tr '\n' ' ')"; tail -2 gpurun_out/step02_r2f.err
ncu --set full --clock-control none --import-source on -k regex:regrid_walk -c 1 -o gpurun_out/prof_regrid_r2d -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2d.log 2>&1; tail -1 gpurun_out/ncu_regrid_r2d.log | cut -c1-200
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2f.log 2> gpurun_out/bench_r2f.err; tail -c 2500 gpurun_out/bench_r2f.log; tail -3 gpurun_out/bench_r2f.err
ncu --set full --clock-control none --import-source on -k regex:pgw_column -s 3 -c 1 -o gpurun_out/prof_column_r2 -f python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 0 --streams 1 > gpurun_out/ncu_column_r2.log 2>&1; tail -1 gpurun_out/ncu_column_r2.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_launches_r2.log 2>&1; wc -l gpurun_out/launches_r2.csv
