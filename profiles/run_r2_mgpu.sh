#!/bin/bash
# multi-GPU call: slab parity (tests/mgpu_check.py) and the bench line at N ranks; usage: run_r2_mgpu.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py 256 16 2>&1 | grep -E "PASS|FAIL|MGPU|hist|Error|error" | cut -c1-300 | tail -40
timeout 600 $TR --master-port 29512 tests/mgpu_check.py 1024 0 2>&1 | grep -E "PASS|FAIL|MGPU|hist|Error|error" | cut -c1-300 | tail -40
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu.json; tail -5 gpurun_out/r2_bench_${N}gpu.err
MPBP_PUSH_FUSED=0 timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_nofusedpush.json 2> gpurun_out/r2_bench_${N}gpu_nofusedpush.err; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu_nofusedpush.json; tail -3 gpurun_out/r2_bench_${N}gpu_nofusedpush.err
MPBP_NCCL_ALLREDUCE=1 timeout 900 $TR --master-port 29515 bench.py --gpus $N --steps 40 --warmup 3 --no-parity > gpurun_out/r2_bench_${N}gpu_ncclallreduce.json 2> gpurun_out/r2_bench_${N}gpu_ncclallreduce.err; echo rc=$?
cat gpurun_out/r2_bench_${N}gpu_ncclallreduce.json; tail -3 gpurun_out/r2_bench_${N}gpu_ncclallreduce.err
