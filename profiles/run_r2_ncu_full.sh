#!/bin/bash
# ncu --set full with source correlation for the level-0 kernels of the apply (eager launches, MPBP_GRAPH=0):
# sweep, A.x, pre-smoothing pair, residual+restriction (launch 66..69), prolongation+sweep, sweep+Chebyshev, residual (106..108)
set -x
mkdir -p gpurun_out
export MPBP_GRAPH=0
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_stokes_x' -s 66 -c 4 -o gpurun_out/r2_full_a -f python profiles/prof_kernels.py > gpurun_out/ncu_full_a.log 2>&1
echo rc=$?
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_stokes_x' -s 106 -c 3 -o gpurun_out/r2_full_b -f python profiles/prof_kernels.py > gpurun_out/ncu_full_b.log 2>&1
echo rc=$?
python profiles/ncu_table.py gpurun_out/r2_full_a.ncu-rep gpurun_out/r2_full_b.ncu-rep | tee gpurun_out/r2_ncu_full_table.txt
ls -la gpurun_out/*.ncu-rep
