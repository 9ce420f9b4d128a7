#!/bin/bash
# round 2, first GPU call: parity (incl. the large-size tests), smoke, bench line, fusion/knob sweep, solve configs
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "passed|failed|error|Error|\[large|\[hist|assert" | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo rc=$?; cat gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 600 python profiles/fuse_sweep.py 4096 2>&1 | tail -12
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_a.txt 2>&1; cat gpurun_out/r2_kernel_table_a.txt
timeout 900 python profiles/solve_configs.py 2>&1 | tail -10
