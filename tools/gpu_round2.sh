#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_hostcpp.py -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
tail -3 gpurun_out/gpu_tests.log
grep -q " failed\| error" gpurun_out/gpu_tests.log && { grep -B60 "short test summary" gpurun_out/gpu_tests.log | tail -90; exit 1; }
python bench.py > gpurun_out/bench_r02_n200.json 2> gpurun_out/bench_r02_n200.err; tail -3 gpurun_out/bench_r02_n200.err
cut -c1-150 gpurun_out/bench_r02_n200.json
