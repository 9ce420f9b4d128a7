#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for f in 0 1 2 3; do echo "=== FUSE=$f"; MPBP_FUSE=$f python profiles/kernel_table.py 4096 2>&1 | grep -E "vcycle_F|precond"; done
