#!/bin/bash
# Round 2, session 2 (4 GPUs): push variants of the in-kernel rollout transport at the driver's settings
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
P=29800
for V in now deferred kernel; do
  P=$((P+1))
  UPKIE_BENCH_PUSH=$V timeout 200 $TR --master-port $P bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/n4_${V}_20.json 2> $O/n4_${V}_20.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/n4_${V}_20.json").read().strip().splitlines()[-1])
    print("N=4 $V steps 20:", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], (d["config"].get("gather") or {}).get("transport"))
except Exception as e: print("$V failed", e)
PY
done
