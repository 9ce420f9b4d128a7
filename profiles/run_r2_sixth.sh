#!/bin/bash
# N-GPU: apply timeline, bench line with parity, two knob variants
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 profiles/trace_apply.py 4096 ${N}gpu_b 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|_warn_once\|Profiler clears" | head -48
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "jac_ms", d["roofline"]["ms_per_launch"], "apply_ms", d["kernels"]["precond_apply"]["ms"], "Ax_ms", d["kernels"]["apply_A"]["ms"], "parity", d.get("parity"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
P
}
timeout 500 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_${N}gpu_b.json 2> gpurun_out/r2_bench_${N}gpu_b.err; echo rc=$?; show gpurun_out/r2_bench_${N}gpu_b.json; tail -2 gpurun_out/r2_bench_${N}gpu_b.err
MPBP_DIST_MIN_N=512 timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_b_dmin512.json 2>/dev/null; show gpurun_out/r2_bench_${N}gpu_b_dmin512.json
MPBP_WAVE=1 timeout 300 $TR --master-port 29515 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_b_wave1.json 2>/dev/null; show gpurun_out/r2_bench_${N}gpu_b_wave1.json
