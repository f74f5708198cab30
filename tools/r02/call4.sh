#!/bin/bash
# Round-2 GPU call 4 (1 GPU): full GPU test-suite on the limits-on default, driver-like bench line, steady-state ncu capture.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu4.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench4_driver.json 2> $O/bench4_driver.err
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/bench4_300.json 2> $O/bench4_300.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench4_reference.json 2> $O/bench4_reference.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 200 -c 1 -f -o $O/prof4_steady python bench.py --steps 210 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/ncu4_steady.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches4.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu4_launches.log 2>&1
tail -15 $O/pytest_gpu4.log
cat $O/bench4_driver.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('driver-like', '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['e2e'].get('median_ms_per_step'), d['clocks'])
print(d['cpu_baseline']); print(d.get('other_workloads')); print(d['roofline'])"
python -c "
import json
d=json.loads(open('$O/bench4_300.json').read().strip().splitlines()[-1]); print('300', '%.4g'%d['value'], d['roofline']['kernel_ms'], 'e2e %.4g'%d['e2e']['value'])
d=json.loads(open('$O/bench4_reference.json').read().strip().splitlines()[-1]); print('ref', '%.4g'%d['value'], d['ms_per_step'], d['cpu_baseline'])"
cat gpurun_out/parity_report.json
