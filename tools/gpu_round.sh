#!/bin/bash
# One single-GPU round trip: tests, then the A/B probes of the ELL family.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
tail -4 gpurun_out/gpu_tests.log
grep -q " failed\| error" gpurun_out/gpu_tests.log && { grep -B30 "short test summary" gpurun_out/gpu_tests.log | tail -60; exit 1; }
timeout 200 python tools/option_probe.py 200 ell_coded=0,1 fuse_p=0,1 > gpurun_out/probe_200.log 2>&1
grep "^{" gpurun_out/probe_200.log
timeout 120 python tools/option_probe.py 100 ell_coded=0,1 fuse_p=0,1 > gpurun_out/probe_100.log 2>&1
grep "^{" gpurun_out/probe_100.log
timeout 120 python tools/ell_probe.py 200 > gpurun_out/ell_probe.log 2>&1
grep "^{" gpurun_out/ell_probe.log
