#!/bin/bash
# LL halo protocol check at N ranks: slab parity, apply timeline, bench line
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py 256 16 2>&1 | grep -E "PASS|FAIL|MGPU|hist|Error|error" | cut -c1-200 | tail -25
timeout 300 $TR --master-port 29512 tests/mgpu_check.py 1024 0 2>&1 | grep -E "PASS|FAIL|MGPU|hist|Error|error" | cut -c1-200 | tail -25
timeout 300 $TR --master-port 29531 profiles/trace_apply.py 4096 ${N}gpu${2} 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|_warn_once\|Profiler clears" | head -40
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu${2}.json 2> gpurun_out/r2_bench_${N}gpu${2}.err; echo rc=$?
python - <<P
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu${2}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["kernels"]["precond_apply"]["ms"], d["kernels"]["apply_A"]["ms"])
P
tail -3 gpurun_out/r2_bench_${N}gpu${2}.err
