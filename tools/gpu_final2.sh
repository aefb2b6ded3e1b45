#!/bin/bash
# Round-end single-GPU pass after the ILU / IC / IRILU / Multigrid work: smoke, the whole -m gpu suite, the
# bench line, ncu captures of the new kernels, compute-sanitizer over their tests (time-bounded).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/final2_smoke.log
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/final2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/final2_gpu_tests.log
timeout 420 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/final2_bench.json
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:'k_trisolve_sf_short|k_tri_ir|k_mg_spmv|k_mg_jacobi_update|k_mg_restrict' -c 12 -o gpurun_out/r02_prof_precond_n100 -f python tools/ncu_target3.py 100 > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_trifactor.py tests/test_gpu_multigrid.py -m gpu -x -q -k "bit_exact and not unstructured" > gpurun_out/sanitizer_memcheck_precond.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_precond.log
