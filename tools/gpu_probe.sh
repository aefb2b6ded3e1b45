#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -x -q -k "fused_into or cg_pressure" > gpurun_out/gpu_tests.log 2>&1
tail -2 gpurun_out/gpu_tests.log
for n in 200 100; do
  timeout 400 python tools/option_probe.py $n ell_tma=1 fuse_p=0,1 tma_stages=2,3 ell_minb=3,4 > gpurun_out/probe_cgptma_$n.log 2>&1
  grep "^{" gpurun_out/probe_cgptma_$n.log | cut -c1-100,190-330 || tail -5 gpurun_out/probe_cgptma_$n.log
done
