#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_ell' -s 6 -c 1 \
   -o gpurun_out/r02_prof_ellpipe_n200 -f python tools/ncu_target.py 200 fuse_p=0 > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log
