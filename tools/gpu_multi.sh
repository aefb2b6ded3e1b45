#!/bin/bash
# usage: gpu_multi.sh NGPUS [cells...] -- all GPU tests (incl. the decomposed path), then A/B of the CG loop options
set -u
N=$1; shift
CELLS=${@:-100 200}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/mg${N}_tests.log 2>&1
tail -3 gpurun_out/mg${N}_tests.log
grep -q " failed\| error" gpurun_out/mg${N}_tests.log && { tail -60 gpurun_out/mg${N}_tests.log; exit 1; }
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for n in $CELLS; do
  run tools/option_probe.py $n fuse_p=0,1 > gpurun_out/mg${N}_probe_$n.log 2>&1
  grep "^{" gpurun_out/mg${N}_probe_$n.log || tail -20 gpurun_out/mg${N}_probe_$n.log
done
for fp in 0 1; do
  run tools/trace_iter.py 100 fused_pcg=0 fuse_p=$fp > gpurun_out/mg${N}_trace_100_fp$fp.log 2>&1
  grep "^{" gpurun_out/mg${N}_trace_100_fp$fp.log | head -1 || tail -20 gpurun_out/mg${N}_trace_100_fp$fp.log
done
