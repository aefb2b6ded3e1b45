#!/bin/bash
set -u
N=$1
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run tools/option_probe.py 100 fuse_p=0,1 > gpurun_out/mg${N}_probe_100.log 2>&1
grep "^{" gpurun_out/mg${N}_probe_100.log || tail -20 gpurun_out/mg${N}_probe_100.log
for fp in 0 1; do
  run tools/trace_iter.py 100 fused_pcg=0 fuse_p=$fp > gpurun_out/mg${N}_trace_100_fp$fp.log 2>&1
  grep "^{" gpurun_out/mg${N}_trace_100_fp$fp.log || tail -20 gpurun_out/mg${N}_trace_100_fp$fp.log
done
