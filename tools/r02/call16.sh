#!/bin/bash
# block size (barrier coupling of the warps of a block) x residual threshold on the headline kernel
mkdir -p gpurun_out/r02e
for blk in 32 64 96 128 160; do
  for thr in 1e-7; do
  UPKIE_B200_HOST_BLOCK=$blk UPKIE_BENCH_RESIDUAL_THRESHOLD=$thr UPKIE_BENCH_DEVICE_ONLY=1 timeout 300 python bench.py --steps 200 --warmup 100 --no-cpu-baseline --no-other-workloads > gpurun_out/r02e/blk_${blk}_$thr.json 2> gpurun_out/r02e/blk_${blk}_$thr.err
  python -c "
import json; j=json.loads(open('gpurun_out/r02e/blk_${blk}_$thr.json').read().strip().splitlines()[-1]); print('block $blk thr $thr: ms %.4f kernel_ms %.4f' % (j['ms_per_step'], j['roofline']['kernel_ms']))"
  done
done
