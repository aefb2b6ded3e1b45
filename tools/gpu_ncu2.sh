#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:'k_bicg_step|k_spmv_ell' -s 20 -c 5 -o gpurun_out/r02_prof_bicgstab_n200 -f python tools/ncu_target2.py bicgstab > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 400 $N -k regex:'k_cg_xr_bj' -s 3 -c 1 -o gpurun_out/r02_prof_bj4_n200 -f python tools/ncu_target2.py bj4 > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
timeout 400 $N -k regex:'k_gmres_mgs|k_gmres_normalize' -s 600 -c 3 -o gpurun_out/r02_prof_gmres_channel -f python tools/ncu_target2.py gmres > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -x -q -k "ell or fused_into or isai or gisai or block_jacobi or cg_pressure" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -k "ell_pattern" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
