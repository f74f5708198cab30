#!/bin/bash
# Round-2 GPU call 8 (1 GPU): the final step kernel (friction impulses in the state row): timing, steady-state capture, launch list
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/bench8_300.json 2> $O/bench8_300.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 200 -c 1 -f -o $O/prof8_steady python bench.py --steps 210 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/ncu8_steady.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 10 -c 1 -f -o $O/prof8_early python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu8_early.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches8.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu8_launches.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu8.log
tail -3 $O/pytest_gpu8.log
python -c "
import json
d=json.loads(open('$O/bench8_300.json').read().strip().splitlines()[-1]); print('300', '%.4g'%d['value'], d['roofline']['kernel_ms'], 'e2e %.4g'%d['e2e']['value'])"
