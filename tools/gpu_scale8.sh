#!/bin/bash
# 8-GPU box: the 4- and 8-rank decomposed parity tests, then the weak-scaling curve at 200^3 per GPU
# (N = 1, 2, 4, 8 on the SAME box), the device timeline at N = 8, and 100^3 per GPU at N = 8
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "procs3-False or procs4-False" > gpurun_out/mg8_tests.log 2>&1
tail -3 gpurun_out/mg8_tests.log
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
python bench.py --no-extra > gpurun_out/scale_r02_1gpu_n200.json 2> gpurun_out/scale_r02_1gpu_n200.err
grep "^{" gpurun_out/scale_r02_1gpu_n200.json | cut -c1-120
for N in 2 4 8; do
  run $N bench.py --gpus $N > gpurun_out/scale_r02_${N}gpu_n200.json 2> gpurun_out/scale_r02_${N}gpu_n200.err
  grep "^{" gpurun_out/scale_r02_${N}gpu_n200.json | cut -c1-120; tail -1 gpurun_out/scale_r02_${N}gpu_n200.err | cut -c1-200
done
run 8 tools/trace_iter.py 200 fused_pcg=0 > gpurun_out/mg8_trace_200.log 2>&1
grep "^{" gpurun_out/mg8_trace_200.log | head -1 | cut -c1-900
run 8 bench.py --gpus 8 --cells 100 > gpurun_out/scale_r02_8gpu_n100.json 2> gpurun_out/scale_r02_8gpu_n100.err
grep "^{" gpurun_out/scale_r02_8gpu_n100.json | cut -c1-120
