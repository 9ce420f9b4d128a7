#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for mb in 0 6 7; do echo "=== JAC_MINB=$mb"; MPBP_JAC_MINB=$mb python profiles/kernel_table.py 4096 2>&1 | grep -E "k_stokes|jacobi_P|k_div|k_grad|vcycle|precond"; done
