#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|large parity" | cut -c1-420 > gpurun_out/r2_gpu_tests.txt; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gpu_tests.txt | tail -8
