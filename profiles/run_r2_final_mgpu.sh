#!/bin/bash
# Final N-GPU evidence of round 2: slab parity (tests/mgpu_check.py), the bench line with its parity block, the
# 8192^2 apply + SpMV sweep (BASELINE configs[4]) and rank 0's apply timeline.  usage: run_r2_final_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/mgpu_check.py 1024 0 > gpurun_out/r2_mgpu_parity_${N}gpu.log 2>&1; echo parity rc=$?
grep -E "PASS|FAIL|MGPU|hist|deviates" gpurun_out/r2_mgpu_parity_${N}gpu.log | cut -c1-230 | tail -32
timeout 500 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_${N}gpu_final.json 2> gpurun_out/r2_bench_${N}gpu_final.err; echo bench rc=$?
cat gpurun_out/r2_bench_${N}gpu_final.json; tail -2 gpurun_out/r2_bench_${N}gpu_final.err
timeout 400 $TR --master-port 29526 bench.py --gpus $N --workload apply8192 --steps 5 --warmup 3 > gpurun_out/r2_apply8192_${N}gpu.json 2> gpurun_out/r2_apply8192_${N}gpu.err; echo sweep rc=$?
cat gpurun_out/r2_apply8192_${N}gpu.json; tail -2 gpurun_out/r2_apply8192_${N}gpu.err
timeout 300 $TR --master-port 29531 profiles/trace_apply.py 4096 ${N}gpu_final > /dev/null 2>&1; head -12 gpurun_out/trace_apply_${N}gpu_final.txt
