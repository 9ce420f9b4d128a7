#!/bin/bash
# 8-GPU call (charged 8x): slab parity, the bench line, two knobs, the 8192^2 throughput sweep
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tests/mgpu_check.py 1024 0 2>&1 | grep -E "PASS|FAIL|MGPU|hist|Error|error" | cut -c1-260 | tail -30
timeout 400 $TR --master-port 29522 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
MPBP_COARSE=64 timeout 400 $TR --master-port 29523 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_coarse64.json 2>/dev/null; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu_coarse64.json
MPBP_DIST_MIN_N=512 timeout 400 $TR --master-port 29524 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_dmin512.json 2>/dev/null; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu_dmin512.json
MPBP_PUSH_FUSED=0 MPBP_NCCL_ALLREDUCE=1 MPBP_ORTH=mgs timeout 400 $TR --master-port 29525 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_r1path.json 2>/dev/null; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu_r1path.json
timeout 400 $TR --master-port 29526 bench.py --gpus $N --workload apply8192 --steps 5 --warmup 3 > gpurun_out/r2_apply8192_${N}gpu.json 2> gpurun_out/r2_apply8192_${N}gpu.err; echo rc=$?
cat gpurun_out/r2_apply8192_${N}gpu.json; tail -3 gpurun_out/r2_apply8192_${N}gpu.err
