#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
tail -4 gpurun_out/gpu_tests.log
grep -q " failed\| error" gpurun_out/gpu_tests.log && { grep -B40 "short test summary" gpurun_out/gpu_tests.log | tail -70; exit 1; }
for n in 200 100; do
  timeout 400 python tools/option_probe.py $n fuse_p=0 ell_minb=3,4 ell_chunk=1,2 > gpurun_out/probe_a_$n.log 2>&1
  grep "^{" gpurun_out/probe_a_$n.log || tail -5 gpurun_out/probe_a_$n.log
  timeout 400 python tools/option_probe.py $n fuse_p=1 ell_minb_cgp=2,3,4 ell_chunk=1,2 > gpurun_out/probe_b_$n.log 2>&1
  grep "^{" gpurun_out/probe_b_$n.log || tail -5 gpurun_out/probe_b_$n.log
done
