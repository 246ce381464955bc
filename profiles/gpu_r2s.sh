#!/bin/bash
# round 2, GPU call s: does the smoothing rate depend on the row stride (2 MB pages)?  K = 19 / 8 / 4 plevs
cd "$(dirname "$0")/.."
for k in 19 8 4; do
  python tests/bench_step02.py --reps 7 --plevs $k > gpurun_out/step02_r2s_k$k.json 2> gpurun_out/step02_r2s_k$k.err
  echo "== plevs $k $(grep -o '"smoothing": {"ms": [0-9.]*, "algorithmic_bytes": [0-9]*, "achieved_gbs": [0-9.]*' gpurun_out/step02_r2s_k$k.json)"; tail -1 gpurun_out/step02_r2s_k$k.err | cut -c1-200
done
