#!/bin/bash
mkdir -p gpurun_out/r02f
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_mpc_step --launch-skip 60 -c 4 --csv --log-file gpurun_out/r02f/mpc_ncu.csv python bench.py --workload mpc --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02f/mpc_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r02f/mpc_ncu.csv | cut -d, -f5,12-15 | tail -9
