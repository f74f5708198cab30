#!/bin/bash
mkdir -p gpurun_out/r02f
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mpc_step --launch-skip 60 -c 1 -f -o gpurun_out/r02f/prof_mpc python bench.py --workload mpc --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02f/mpc_under_ncu2.log 2>&1
ls -la gpurun_out/r02f/
