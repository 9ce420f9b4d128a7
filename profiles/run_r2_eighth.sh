#!/bin/bash
# N-GPU: apply timeline, bench default + knob variants (no parity)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 profiles/trace_apply.py 4096 ${N}gpu_c 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|_warn_once\|Profiler clears" | head -30
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "jac_ms", d["roofline"]["ms_per_launch"], "apply_ms", d["kernels"]["precond_apply"]["ms"], "Ax_ms", d["kernels"]["apply_A"]["ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
P
}
run() { tag=$1; shift; env "$@" timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_c_$tag.json 2>/dev/null; show gpurun_out/r2_bench_${N}gpu_c_$tag.json; }
run default X=1
run pf6 MPBP_PF=6
run rs14 MPBP_RS=14
run dmin512 MPBP_DIST_MIN_N=512
