#!/bin/bash
# sensitivity of the headline kernel to the residual threshold (number of PGS sweeps)
mkdir -p gpurun_out/r02e
for thr in 0 1e-9 1e-7 1e-5 1e-3; do
  UPKIE_BENCH_RESIDUAL_THRESHOLD=$thr UPKIE_BENCH_DEVICE_ONLY=1 timeout 300 python bench.py --steps 200 --warmup 100 --no-cpu-baseline --no-other-workloads > gpurun_out/r02e/thr_$thr.json 2> gpurun_out/r02e/thr_$thr.err
  python -c "
import json; j=json.loads(open('gpurun_out/r02e/thr_$thr.json').read().strip().splitlines()[-1]); print('thr $thr: ms %.4f kernel_ms %.4f' % (j['ms_per_step'], j['roofline']['kernel_ms']))"
done
