#!/bin/bash
# usage: gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 256 16 > gpurun_out/mgpu$N.log 2>&1
grep -E "FAIL|MGPU|   eta" gpurun_out/mgpu$N.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_check.py 2048 > gpurun_out/mgpu${N}b.log 2>&1
grep -E "FAIL|MGPU|   eta" gpurun_out/mgpu${N}b.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo rc=$?; tail -n 5 gpurun_out/bench_$N.err; cat gpurun_out/bench_$N.json
