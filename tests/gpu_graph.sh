#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for g in 0 1; do echo "=== GRAPH=$g"; MPBP_GRAPH=$g python profiles/kernel_table.py 4096 2>&1 | grep -E "jacobi_F|vcycle|precond"; done
python bench.py --steps 40 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-300
