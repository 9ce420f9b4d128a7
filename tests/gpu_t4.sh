#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for f in 1 0; do echo "=== FUSED_MGS=$f"; MPBP_FUSED_MGS=$f python bench.py --steps 40 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels']['precond_apply']['ms'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; done
