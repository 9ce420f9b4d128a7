#!/bin/bash
# ncu --set full with source correlation for the level-0 kernels (eager launches, MPBP_GRAPH=0).  prof_kernels.py
# brackets the interesting launches with cudaProfilerStart/Stop, so plan creation is not counted (--profile-from-start
# off).  k_stokes_x launches after the start: 0-2 sweeps, 3 A.x, 4 pre-smoothing pair, 5 residual+restriction,
# 6..17 levels 1-3, 18 prolongation+sweep, 19 sweep+Chebyshev, 20 residual.
mkdir -p gpurun_out
export MPBP_GRAPH=0
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1; echo plain rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_stokes_x' -s 2 -c 4 -o gpurun_out/r2_full_a -f python profiles/prof_kernels.py > gpurun_out/ncu_full_a.log 2>&1
echo full_a rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_stokes_x' -s 18 -c 3 -o gpurun_out/r2_full_b -f python profiles/prof_kernels.py > gpurun_out/ncu_full_b.log 2>&1
echo full_b rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_poisson|k_div|k_grad' -c 3 -o gpurun_out/r2_full_c -f python profiles/prof_kernels.py > gpurun_out/ncu_full_c.log 2>&1
echo full_c rc=$?
python profiles/ncu_table.py gpurun_out/r2_full_a.ncu-rep gpurun_out/r2_full_b.ncu-rep gpurun_out/r2_full_c.ncu-rep --traffic gpurun_out/r2_ncu_traffic.json | tee gpurun_out/r2_ncu_full_table.txt
cat gpurun_out/r2_ncu_traffic.json
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full_a.log
