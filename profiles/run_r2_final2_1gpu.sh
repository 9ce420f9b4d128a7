#!/bin/bash
# 1-GPU re-measurement after the fused pressure cycle: GPU tests (with the measured parity lines), bench line, kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|large parity" | cut -c1-420 > gpurun_out/r2_gpu_tests.txt; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gpu_tests.txt | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; echo bench rc=$?; cat gpurun_out/r2_bench_1gpu_final.json
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_final.txt 2>&1; tail -4 gpurun_out/r2_kernel_table_final.txt
timeout 300 python profiles/trace_apply.py 4096 1gpu_final > /dev/null 2>&1; head -3 gpurun_out/trace_apply_1gpu_final.txt
