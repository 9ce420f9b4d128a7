#!/bin/bash
for pf in 0 2 3 4 5 6; do
echo "=== PF=$pf"; MPBP_PF=$pf python profiles/kernel_table.py 4096 2>&1 | grep -E "k_stokes|jacobi_P|vcycle|precond"
done
