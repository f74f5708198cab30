#!/bin/bash
mkdir -p gpurun_out/r02e
for i in 1 2; do
UPKIE_BENCH_DEVICE_ONLY=1 timeout 300 python bench.py --steps 200 --warmup 100 --no-cpu-baseline --no-other-workloads > gpurun_out/r02e/pair_$i.json 2> gpurun_out/r02e/pair_$i.err
python -c "
import json; j=json.loads(open('gpurun_out/r02e/pair_$i.json').read().strip().splitlines()[-1]); print('run $i: ms %.4f kernel_ms %.4f' % (j['ms_per_step'], j['roofline']['kernel_ms']))"
done
UPKIE_BENCH_DEVICE_ONLY=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > gpurun_out/r02e/pair_drv.json 2> gpurun_out/r02e/pair_drv.err
python -c "
import json; j=json.loads(open('gpurun_out/r02e/pair_drv.json').read().strip().splitlines()[-1]); print('driver-style: ms %.4f kernel_ms %.4f' % (j['ms_per_step'], j['roofline']['kernel_ms']))"
