#!/bin/bash
# usage: run_multi_gpu.sh N [check]  -- bounded multi-GPU run: optional parity check at n=2048, then the bench
# experimental paths are selected through the environment, e.g.  MPBP_PUSH_FUSED=1 MPBP_COARSE=64 run_multi_gpu.sh 2 check
N=$1
if [ "$2" = "check" ]; then
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_check.py 2048 > gpurun_out/mgpu${N}b.log 2>&1
echo check_rc=$?
grep -E "FAIL|MGPU|   eta" gpurun_out/mgpu${N}b.log
fi
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo bench_rc=$?; tail -n 3 gpurun_out/bench_$N.err; cat gpurun_out/bench_$N.json
