#!/bin/bash
# 1-GPU validation of the fused pressure-Poisson kernels and the coarse-row prefetch: tests, apply timings per knob, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|large parity" | cut -c1-420 > gpurun_out/r2_gpu_tests.txt; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gpu_tests.txt | tail -8
for kn in "X=1" "MPBP_PFC=0" "MPBP_FUSE=7" "MPBP_FUSE=7 MPBP_PFC=0"; do echo "== $kn"; env $kn timeout 300 python profiles/trace_apply.py 4096 "1gpu_$(echo $kn | tr ' =' '__')" 2>&1 | grep "apply .* ms without"; done
timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_e.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "jac_ms", d["roofline"]["ms_per_launch"], "apply", d["kernels"]["precond_apply"], "Ax_ms", d["kernels"]["apply_A"]["ms"], "parity", d.get("parity"))
P
timeout 600 python profiles/kernel_table.py 4096 2>&1 | tail -15
