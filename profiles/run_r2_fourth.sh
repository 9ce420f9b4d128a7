#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-600 | tail -20
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo rc=$?; cat gpurun_out/r2_bench_b.json; tail -3 gpurun_out/r2_bench_b.err
MPBP_ORTH=mgs timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu --no-parity > gpurun_out/r2_bench_b_mgs.json 2>/dev/null; cat gpurun_out/r2_bench_b_mgs.json
timeout 900 python profiles/fuse_sweep.py 4096 2>&1 | tail -14
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_c.txt 2>&1; cat gpurun_out/r2_kernel_table_c.txt
export MPBP_GRAPH=0
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats --section ComputeWorkloadAnalysis"
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu $SEC --clock-control none -k regex:'k_stokes_x' -s 64 -c 48 -o gpurun_out/r2_prof_stokes -f python profiles/prof_kernels.py > gpurun_out/ncu_a.log 2>&1
echo rc=$?
python profiles/ncu_table.py gpurun_out/r2_prof_stokes.ncu-rep > gpurun_out/r2_ncu_stokes_table.txt 2>&1; cat gpurun_out/r2_ncu_stokes_table.txt
for f in gpurun_out/*.ncu-rep; do sz=$(stat -c %s $f); if [ $sz -gt 20000000 ]; then rm -f $f; fi; done
