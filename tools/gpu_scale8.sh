#!/bin/bash
# 8-GPU box: the 8-rank decomposed parity test, the bench at N = 8 and 4 (200^3 per GPU), device timeline
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "procs4-False" > gpurun_out/mg8_tests.log 2>&1
tail -3 gpurun_out/mg8_tests.log
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for N in 8 4; do
  run $N bench.py --gpus $N > gpurun_out/bench_r02_${N}gpu_n200.json 2> gpurun_out/bench_r02_${N}gpu_n200.err
  cut -c1-300 gpurun_out/bench_r02_${N}gpu_n200.json; tail -2 gpurun_out/bench_r02_${N}gpu_n200.err
done
run 8 tools/trace_iter.py 200 fused_pcg=0 > gpurun_out/mg8_trace_200.log 2>&1
grep "^{" gpurun_out/mg8_trace_200.log | head -2
run 8 bench.py --gpus 8 --cells 100 > gpurun_out/bench_r02_8gpu_n100.json 2> gpurun_out/bench_r02_8gpu_n100.err
cut -c1-300 gpurun_out/bench_r02_8gpu_n100.json
