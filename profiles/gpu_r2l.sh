#!/bin/bash
# round 2, GPU call l (2+ GPUs): the latitude-band exchange fused over peer memory
cd "$(dirname "$0")/.."
N=${1:-2}
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for x in auto nccl; do
  $TR --master-port 29521 tests/multigpu_latband.py --exchange $x > gpurun_out/latband_parity_${x}_n$N.log 2>&1; tail -2 gpurun_out/latband_parity_${x}_n$N.log | cut -c1-300
done
$TR --master-port 29522 tests/multigpu_latband.py --config5 --exchange auto > gpurun_out/latband_parity_c5_n$N.log 2>&1; tail -1 gpurun_out/latband_parity_c5_n$N.log | cut -c1-300
$TR --master-port 29524 bench.py --gpus $N --steps 20 --warmup 5 ${2:-} > gpurun_out/bench_l_n$N.log 2> gpurun_out/bench_l_n$N.err; python - <<PY
import json
for l in open('gpurun_out/bench_l_n$N.log'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['config']['latband'])[:1800]); print(d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value'])
PY
tail -5 gpurun_out/bench_l_n$N.err
