#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small cases): memcheck, synccheck, racecheck
set -u
mkdir -p gpurun_out
K='tma or ell or fused_into or isai or gisai or block_jacobi or cg_pressure or gmres or bicgstab_momentum or histogram'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -x -q \
   -k "($K) and not 200_cubed and not 100_cubed and not half_million and not unstructured" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -x -q \
   -k "tma or test_gmres or ell_pattern" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/sanitizer_synccheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q \
   -k "tma or ell_pattern" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
