#!/bin/bash
# round 2, third GPU call: full parity suite, then ncu on the first V-cycle of the apply (eager launches), summarised
# on the box (only text comes back; the .ncu-rep files are dropped if they are large)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|\[large|\[hist|Error" | cut -c1-700 | tail -70
timeout 600 python profiles/fuse_sweep.py 4096 2>&1 | tail -12
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_b.txt 2>&1; cat gpurun_out/r2_kernel_table_b.txt
export MPBP_GRAPH=0
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats --section ComputeWorkloadAnalysis"
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu $SEC --clock-control none -k regex:'k_stokes_x' -s 68 -c 44 -o gpurun_out/r2_prof_stokes -f python profiles/prof_kernels.py > gpurun_out/ncu_a.log 2>&1
echo rc=$?
python profiles/ncu_table.py gpurun_out/r2_prof_stokes.ncu-rep > gpurun_out/r2_ncu_stokes_table.txt 2>&1; cat gpurun_out/r2_ncu_stokes_table.txt
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu $SEC --clock-control none -k regex:'k_poisson|k_div|k_grad|k_restrict_P|k_prolong_add_P|k_combine|k_jacobi0' -s 20 -c 40 -o gpurun_out/r2_prof_light -f python profiles/prof_kernels.py > gpurun_out/ncu_b.log 2>&1
echo rc=$?
python profiles/ncu_table.py gpurun_out/r2_prof_light.ncu-rep > gpurun_out/r2_ncu_light_table.txt 2>&1; cat gpurun_out/r2_ncu_light_table.txt
ls -la gpurun_out/
for f in gpurun_out/*.ncu-rep; do sz=$(stat -c %s $f); if [ $sz -gt 20000000 ]; then rm -f $f; fi; done
tail -n 3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
