#!/bin/bash
# Round-2 GPU call 1: GPU tests, joint-limit solver timing (0 off / 1 slow path / 2 ten-row / 3 hybrid), ncu captures.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for L in 0 1 2 3; do
  UPKIE_BENCH_JOINT_LIMITS=$L timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_limits$L.json 2> $O/bench_limits$L.err
done
for L in 0 3; do
  UPKIE_BENCH_JOINT_LIMITS=$L timeout 300 python bench.py --workload pendulum --steps 1000 --warmup 100 --no-cpu-baseline > $O/bench_pendulum_limits$L.json 2> $O/bench_pendulum_limits$L.err
done
timeout 300 python bench.py --workload mpc --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_mpc.json 2> $O/bench_mpc.err
for L in 0 2 3; do
  UPKIE_BENCH_JOINT_LIMITS=$L UPKIE_BENCH_CUDA_PROFILER=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --launch-skip 5 -c 1 -f -o $O/prof_step_limits$L python bench.py --steps 10 --warmup 5 --no-cpu-baseline > $O/ncu_limits$L.log 2>&1
done
UPKIE_BENCH_JOINT_LIMITS=2 UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches_limits2.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_launches.log 2>&1
tail -3 $O/pytest_gpu.log
for L in 0 1 2 3; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_limits$L.json").read().strip().splitlines()[-1])
    print("limits$L", d["value"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("limits$L failed", e)
PY
done
ls -la $O
