#!/bin/bash
# Round-2 GPU call 11 (1 GPU): host-path pipeline sweep with the limits-on kernel (0.11-0.14 ms): chunks, pipelines
mkdir -p gpurun_out/r02
O=gpurun_out/r02
: > $O/e2e_sweep.txt
run() { # label, env assignments...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', 'e2e %.4g'%d['e2e']['value'], 'median ms %.4f'%d['e2e']['median_ms_per_step'], 'value %.4g'%d['value'])" >> $O/e2e_sweep.txt
}
for C in 1 2 3 4 6 8 12; do run "hybrid chunks=$C" UPKIE_B200_HOST_CHUNKS=$C; done
run "hybrid chunks=4 kstreams=2" UPKIE_B200_HOST_CHUNKS=4 UPKIE_B200_HOST_KERNEL_STREAMS=2
run "zero-copy persistent" UPKIE_B200_ZERO_COPY=1
run "staged copies chunks=4" UPKIE_B200_ZERO_COPY=0 UPKIE_B200_HOST_CHUNKS=4
run "hybrid split .3,.3,.25,.15" UPKIE_B200_HOST_SPLIT=0.3,0.3,0.25,0.15
cat $O/e2e_sweep.txt
