#!/bin/bash
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 40 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; cat gpurun_out/bench_final.json
python profiles/solve_configs.py 2>&1 | tail -6
python profiles/kernel_table.py 4096 > gpurun_out/kernel_table.txt 2>&1; cat gpurun_out/kernel_table.txt
python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1b.csv python profiles/prof_kernels.py > gpurun_out/ncu1.log 2>&1
echo rc=$?
python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stokes -s 64 -c 4 -o gpurun_out/prof_stokes_r1b -f python profiles/prof_kernels.py > gpurun_out/ncu2.log 2>&1
echo rc=$?
