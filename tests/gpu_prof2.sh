#!/bin/bash
set -x
python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stokes -s 64 -c 4 -o gpurun_out/prof_stokes_r1 -f python profiles/prof_kernels.py > gpurun_out/ncu2.log 2>&1
echo rc=$?
tail -n 3 gpurun_out/ncu2.log
