#!/bin/bash
# round 2, GPU call p: last whole-suite check of the committed tree + default bench
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests_r2p.log 2>&1; tail -3 gpurun_out/tests_r2p.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2p.log 2> gpurun_out/bench_r2p.err; python - <<PY
import json
for l in open('gpurun_out/bench_r2p.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_link'], d['cpu_baseline']['value'], d['clocks'])
PY
tail -3 gpurun_out/bench_r2p.err
python tests/bench_files.py --files 16 --raw-only --breakdown > gpurun_out/files_eu_r2b.json 2> gpurun_out/files_eu_r2b.err; grep '^{' gpurun_out/files_eu_r2b.json | tail -1 | cut -c1-700
python tests/bench_files.py --files 4 --ny 721 --nx 1440 --raw-only --breakdown > gpurun_out/files_gl_r2b.json 2> gpurun_out/files_gl_r2b.err; grep '^{' gpurun_out/files_gl_r2b.json | tail -1 | cut -c1-700; tail -2 gpurun_out/files_gl_r2b.err
