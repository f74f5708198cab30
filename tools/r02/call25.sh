#!/bin/bash
# gate-only cost of the body-contact kernels: pendulum 4096 off / on; torque workload with the reset height at 0.25 m
# (robots recycle before the torso box, bottom 0.21 m below the base origin, reaches the floor)
mkdir -p gpurun_out/r02j
python tools/r02/body_kernels_cost.py 2>&1 | tail -2
for bc in 0 1; do
UPKIE_BENCH_BODY_CONTACTS=$bc UPKIE_BENCH_MIN_BASE_HEIGHT=0.25 UPKIE_BENCH_DEVICE_ONLY=1 timeout 300 python bench.py --steps 200 --warmup 100 --no-cpu-baseline --no-other-workloads > gpurun_out/r02j/h025_bc$bc.json 2> gpurun_out/r02j/h025_bc$bc.err
python -c "
import json; j=json.loads(open('gpurun_out/r02j/h025_bc$bc.json').read().strip().splitlines()[-1]); print('torque workload, reset height 0.25 m, body_contacts=$bc: ms %.4f kernel_ms %.4f' % (j['ms_per_step'], j['roofline']['kernel_ms']))"
done
