#!/bin/bash
# round 2, GPU call k: whole GPU suite with the report files, default bench (+ reference arm), file pipelines, final step_02 numbers
cd "$(dirname "$0")/.."
export PGW_PARITY_OUT=$PWD/gpurun_out/parity_lines.jsonl PGW_REFDTYPES_OUT=$PWD/gpurun_out/refdtypes_lines.jsonl
rm -f $PGW_PARITY_OUT $PGW_REFDTYPES_OUT
python -m pytest tests -m gpu -q --tb=short > gpurun_out/tests_r2k.log 2>&1; tail -4 gpurun_out/tests_r2k.log
unset PGW_PARITY_OUT PGW_REFDTYPES_OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2k.log 2> gpurun_out/bench_r2k.err; tail -c 3500 gpurun_out/bench_r2k.log; tail -3 gpurun_out/bench_r2k.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2k.log 2> gpurun_out/bench_ref_r2k.err; tail -c 1500 gpurun_out/bench_ref_r2k.log
python tests/bench_step02.py > gpurun_out/step02_r2k.json 2> gpurun_out/step02_r2k.err
echo "== step02 $(grep -o '"smoothing": {"ms": [0-9.]*\|"regridding": {"ms": [0-9.]*\|"frac_of_peak": [0-9.]*' gpurun_out/step02_r2k.json | tr '\n' ' ')"
ncu --set full --clock-control none --import-source on -k regex:regrid_walk -c 1 -o gpurun_out/prof_regrid_r2f -f python tests/bench_step02.py --reps 1 > gpurun_out/ncu_regrid_r2f.log 2>&1; tail -1 gpurun_out/ncu_regrid_r2f.log | cut -c1-200
python tests/bench_files.py --files 16 --raw-only --breakdown > gpurun_out/files_eu_r2.json 2> gpurun_out/files_eu_r2.err; tail -c 1200 gpurun_out/files_eu_r2.json
python tests/bench_files.py --files 3 --ny 721 --nx 1440 --raw-only --breakdown > gpurun_out/files_gl_r2.json 2> gpurun_out/files_gl_r2.err; tail -c 1200 gpurun_out/files_gl_r2.json; tail -2 gpurun_out/files_gl_r2.err
