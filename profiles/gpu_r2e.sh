#!/bin/bash
# round 2, GPU call e (2 GPUs): the multi-rank paths -- bench with broadcast / e2e link bound / latitude bands,
# band-split step_02, concurrent PCIe probe, latband parity script
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tests/multigpu_latband.py > gpurun_out/latband_parity_n$N.log 2>&1; tail -3 gpurun_out/latband_parity_n$N.log | cut -c1-300
$TR --master-port 29512 tests/multigpu_step02.py > gpurun_out/step02_banded_n$N.log 2>&1; tail -2 gpurun_out/step02_banded_n$N.log | cut -c1-900
$TR --master-port 29513 tests/multigpu_pcie.py > gpurun_out/pcie_n$N.log 2>&1; tail -1 gpurun_out/pcie_n$N.log | cut -c1-1500
$TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; tail -c 5000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
