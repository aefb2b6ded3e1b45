#!/bin/bash
# First hardware run of this branch (everything here was written after the round-1 GPU budget
# was spent).  One gpurun call, 1 GPU, ~2 minutes:
#   gpurun --timeout 600 -- 'bash tools/first_run_on_hardware.sh'
# Order: correctness first (stop on the first failure), then the A/B numbers that decide what
# gets merged into main.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/branch_gpu_tests.log 2>&1
tail -3 gpurun_out/branch_gpu_tests.log
grep -q " failed" gpurun_out/branch_gpu_tests.log && exit 1
# static-width ELL vs pipelined CSR: plain / fused SpMV and the PCG iteration
timeout 120 python tools/ell_probe.py 100 200 > gpurun_out/branch_ell_probe.log 2>&1
grep "^{" gpurun_out/branch_ell_probe.log
# p-update fused into the ELL SpMV (two launches per iteration)
timeout 120 python tools/option_probe.py 100 ell_auto=0,1 fuse_p=0,1 > gpurun_out/branch_fuse_p_100.log 2>&1
grep "^{" gpurun_out/branch_fuse_p_100.log
timeout 120 python tools/option_probe.py 200 ell_auto=0,1 fuse_p=0,1 > gpurun_out/branch_fuse_p_200.log 2>&1
grep "^{" gpurun_out/branch_fuse_p_200.log
# whole-solve number and the device timeline with the early first-trip loads
timeout 120 python bench.py > gpurun_out/branch_bench_n100.log 2>&1
grep "^{" gpurun_out/branch_bench_n100.log | cut -c1-400
timeout 60 python tools/trace_iter.py 100 fused_pcg=0 > gpurun_out/branch_trace_100.log 2>&1
grep "^{" gpurun_out/branch_trace_100.log
