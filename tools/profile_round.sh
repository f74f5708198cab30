#!/bin/bash
# Round profile (run on the GPU box through gpurun): bench lines, ncu launch lists and full captures.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.md here.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_servos_$TAG.json 2> gpurun_out/bench_servos_$TAG.err
python bench.py --workload pendulum --steps 1000 --warmup 100 > gpurun_out/bench_pendulum_$TAG.json 2>> gpurun_out/bench_servos_$TAG.err
python bench.py --workload mpc --steps 300 --warmup 20 > gpurun_out/bench_mpc_$TAG.json 2>> gpurun_out/bench_servos_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>> gpurun_out/bench_servos_$TAG.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=e2e timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_e2e_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 5 -c 1 -f -o gpurun_out/prof_step_$TAG python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu3.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=e2e timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 4 -c 1 -f -o gpurun_out/prof_step_host_$TAG python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu4.log 2>&1
ls -la gpurun_out | tail -20
tail -c 600 gpurun_out/bench_servos_$TAG.json
