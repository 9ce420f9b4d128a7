#!/bin/bash
# 1-GPU: GPU test suite, bench line, apply timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | cut -c1-400 | tail -20
timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_d.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "jac_ms", d["roofline"]["ms_per_launch"], "apply", d["kernels"]["precond_apply"], "Ax_ms", d["kernels"]["apply_A"]["ms"], "parity", d.get("parity"))
P
tail -3 gpurun_out/r2_bench_d.err
timeout 300 python profiles/trace_apply.py 4096 1gpu_c 2>&1 | grep -v "_warn_once\|Profiler clears" | head -64
