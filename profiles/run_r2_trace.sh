#!/bin/bash
# timeline trace of the apply at N ranks (rank 0's kernels); usage: run_r2_trace.sh N [with1]
N=${1:-2}
mkdir -p gpurun_out
[ -n "$2" ] && timeout 300 python profiles/trace_apply.py 4096 1gpu 2>&1 | tail -45
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 profiles/trace_apply.py 4096 ${N}gpu 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|_warn_once\|Profiler clears" | tail -70
