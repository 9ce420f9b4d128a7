#!/bin/bash
# 1-GPU: GPU test suite, smoke, bench line (parity + cpu baseline), kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | cut -c1-400 | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo rc=$?; cat gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_d.txt 2>&1; cat gpurun_out/r2_kernel_table_d.txt
