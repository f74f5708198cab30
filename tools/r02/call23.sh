#!/bin/bash
# Round 2, session 2 (2 GPUs): push variants of the in-kernel rollout transport with the final kernels
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29700
for V in kernel deferred now; do
  for S in "20 5" "256 20"; do
    set -- $S; P=$((P+1))
    UPKIE_BENCH_PUSH=$V timeout 300 $TR --master-port $P bench.py --gpus 2 --steps $1 --warmup $2 --no-cpu-baseline --no-other-workloads > $O/push_${V}_$1.json 2> $O/push_${V}_$1.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/push_${V}_$1.json").read().strip().splitlines()[-1])
    print("$V steps $1:", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], (d["config"].get("gather") or {}).get("transport"), "stall", (d["config"].get("gather") or {}).get("sim_stream_stall_ms_total"))
except Exception as e: print("$V $1 failed", e)
PY
  done
done
UPKIE_BENCH_GATHER=nccl timeout 300 $TR --master-port 29790 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/push_nccl_20.json 2> $O/push_nccl_20.err
python - <<PY
import json
d=json.loads(open("$O/push_nccl_20.json").read().strip().splitlines()[-1]); print("nccl steps 20:", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"])
PY
