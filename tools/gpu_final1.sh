#!/bin/bash
# Final single-GPU pass: full GPU suite, bench line, ncu launch list of the bench command, ncu --set full
# of the iteration kernels with the TMA-fed SpMV.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/gpu_tests.log 2>&1
tail -9 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_r02_n200.json 2> gpurun_out/bench_r02_n200.err; cut -c1-200 gpurun_out/bench_r02_n200.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 500 --csv --log-file gpurun_out/r02_launches_bench_n200.csv \
   python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_ell_tma|k_cg_xr|k_cg_p' -s 12 -c 3 \
   -o gpurun_out/r02_prof_pcg3_tma_n200 -f python tools/ncu_target.py 200 > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
