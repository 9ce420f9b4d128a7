#!/bin/bash
set -x
python bench.py --steps 40 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
