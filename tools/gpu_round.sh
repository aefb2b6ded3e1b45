#!/bin/bash
# One single-GPU round trip: the whole GPU test suite (with timings of the slowest tests).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests.log 2>&1
tail -16 gpurun_out/gpu_tests.log
grep -q " failed\| error" gpurun_out/gpu_tests.log && { grep -B60 "short test summary" gpurun_out/gpu_tests.log | tail -90; exit 1; }
exit 0
