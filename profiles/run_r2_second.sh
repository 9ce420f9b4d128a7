#!/bin/bash
# round 2, second GPU call: full parity suite + ncu --set full of the level-0 kernels (eager launches, no graph)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|\[large|\[hist|AssertionError" | tail -60
export MPBP_GRAPH=0
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_stokes_x|k_jacobi0' -s 8 -c 56 -o gpurun_out/r2_prof_stokes -f python profiles/prof_kernels.py > gpurun_out/ncu_a.log 2>&1
echo rc=$?
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_poisson|k_div|k_grad|k_restrict_P|k_prolong_add_P|k_combine|k_mgs|k_multi' -c 40 -o gpurun_out/r2_prof_light -f python profiles/prof_kernels.py > gpurun_out/ncu_b.log 2>&1
echo rc=$?
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
