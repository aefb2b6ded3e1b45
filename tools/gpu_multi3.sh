#!/bin/bash
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/mg${N}_tests.log 2>&1
tail -2 gpurun_out/mg${N}_tests.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for n in 200 100; do
  run tools/option_probe.py $n ell_tma=2 > gpurun_out/mg${N}_probe_$n.log 2>&1
  grep "^{" gpurun_out/mg${N}_probe_$n.log | cut -c1-70,200-330 || tail -20 gpurun_out/mg${N}_probe_$n.log
done
