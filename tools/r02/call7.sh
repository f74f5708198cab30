#!/bin/bash
# Round-2 GPU call 7 (1 GPU): GPU tests incl. spine mode; steady-state + early captures of the final kernel for the sidecar
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu7.log
tail -6 $O/pytest_gpu7.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench7_driver.json 2> $O/bench7_driver.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 200 -c 1 -f -o $O/prof7_steady python bench.py --steps 210 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/ncu7_steady.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches7.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu7_launches.log 2>&1
timeout 300 python bench.py --workload mpc --steps 300 --warmup 20 --no-cpu-baseline > $O/bench7_mpc.json 2> $O/bench7_mpc.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mpc_step --launch-skip 50 -c 1 -f -o $O/prof7_mpc \
  python bench.py --workload mpc --steps 100 --warmup 20 --no-cpu-baseline > $O/ncu7_mpc.log 2>&1
python -c "
import json
d=json.loads(open('$O/bench7_driver.json').read().strip().splitlines()[-1]); print('driver-like', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['clocks']['sm_mhz'], d['cpu_baseline']['value'])"
cat gpurun_out/parity_report.json | head -12
