#!/bin/bash
# the weak-scaling curve at 200^3 per GPU on ONE 8-GPU box (N = 1, 2, 4, 8), nothing else
set -u
mkdir -p gpurun_out
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
python bench.py --no-extra > gpurun_out/scale_r02_1gpu_n200.json 2> gpurun_out/scale_r02_1gpu_n200.err
grep "^{" gpurun_out/scale_r02_1gpu_n200.json | cut -c1-120
for N in 2 4 8; do
  run $N bench.py --gpus $N > gpurun_out/scale_r02_${N}gpu_n200.json 2> gpurun_out/scale_r02_${N}gpu_n200.err
  grep "^{" gpurun_out/scale_r02_${N}gpu_n200.json | cut -c1-120
done
run 8 bench.py --gpus 8 --cells 100 > gpurun_out/scale_r02_8gpu_n100.json 2> gpurun_out/scale_r02_8gpu_n100.err
grep "^{" gpurun_out/scale_r02_8gpu_n100.json | cut -c1-120
