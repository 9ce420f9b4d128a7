#!/bin/bash
# Final 1-GPU evidence of round 2: GPU tests, smoke, bench line, 8192^2 sweep, kernel table, solves at BASELINE's
# sizes, the ncu launch list of bench.py and the --set full captures of the level-0 kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-300 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 40 --warmup 3 > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; echo bench rc=$?; cat gpurun_out/r2_bench_1gpu_final.json
timeout 600 python bench.py --workload apply8192 --steps 5 --warmup 3 > gpurun_out/r2_apply8192_1gpu.json 2> gpurun_out/r2_apply8192_1gpu.err; echo sweep rc=$?; cat gpurun_out/r2_apply8192_1gpu.json
timeout 600 python profiles/kernel_table.py 4096 > gpurun_out/r2_kernel_table_final.txt 2>&1; cat gpurun_out/r2_kernel_table_final.txt
timeout 600 python profiles/solve_configs.py 2>&1 | grep -v Warn | cut -c1-330 | tail -8
timeout 300 python profiles/trace_apply.py 4096 1gpu_final > /dev/null 2>&1; head -3 gpurun_out/trace_apply_1gpu_final.txt
# launch list of the bench command itself (the plan-creation launches are skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2600 -c 2600 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1; echo launches rc=$?
python profiles/launch_shares.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launch_shares.txt 2>&1; head -16 gpurun_out/r2_launch_shares.txt
export MPBP_GRAPH=0
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stokes_x' -s 2 -c 4 -o gpurun_out/r2_full_a -f python profiles/prof_kernels.py > gpurun_out/ncu_full_a.log 2>&1
echo full_a rc=$?
timeout 300 python profiles/prof_kernels.py > gpurun_out/prof_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stokes_x' -s 18 -c 3 -o gpurun_out/r2_full_b -f python profiles/prof_kernels.py > gpurun_out/ncu_full_b.log 2>&1
echo full_b rc=$?
python profiles/ncu_table.py gpurun_out/r2_full_a.ncu-rep gpurun_out/r2_full_b.ncu-rep --traffic gpurun_out/r2_ncu_traffic.json | tee gpurun_out/r2_ncu_full_table.txt
ls -la gpurun_out/*.ncu-rep
