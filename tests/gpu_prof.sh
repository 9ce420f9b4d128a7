#!/bin/bash
set -x
python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python profiles/prof_kernels.py > gpurun_out/ncu1.log 2>&1
echo rc=$?
python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stokes -c 4 -o gpurun_out/prof_stokes_r1 python profiles/prof_kernels.py > gpurun_out/ncu2.log 2>&1
echo rc=$?
tail -3 gpurun_out/prof_plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out
