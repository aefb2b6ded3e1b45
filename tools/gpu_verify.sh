#!/bin/bash
# What the driver does at round end, in one go: smoke, the GPU suite, both bench arms at N = 1.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
python bench.py --impl reference > gpurun_out/bench_ref_n200.json 2> gpurun_out/bench_ref_n200.err; cut -c1-250 gpurun_out/bench_ref_n200.json
python bench.py > gpurun_out/bench_r02_n200.json 2> gpurun_out/bench_r02_n200.err; cut -c1-200 gpurun_out/bench_r02_n200.json
