#!/bin/bash
# Round-2 GPU call 12 (2 GPUs): deferred in-kernel rollout transport validated and timed against the immediate form
mkdir -p gpurun_out/r02
O=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/multicast_check.py > $O/multicast_check2b.log 2>&1; echo "rc=$?" >> $O/multicast_check2b.log
grep -h "rollout\|rc=" $O/multicast_check2b.log | sort | uniq -c
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/b12_n1.json 2> $O/b12_n1.err
P=29800
for V in "deferred multicast" "now multicast" "deferred peerstore"; do
  set -- $V; P=$((P+1))
  UPKIE_BENCH_PUSH=$1 UPKIE_BENCH_GATHER=$2 timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $O/b12_$1_$2.json 2> $O/b12_$1_$2.err
done
for f in n1 deferred_multicast now_multicast deferred_peerstore; do python - <<PY
import json
try:
    d=json.loads(open("$O/b12_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], (d["config"].get("gather") or {}).get("transport"))
except Exception as e: print("$f failed", e)
PY
done
tail -n 4 $O/b12_deferred_multicast.err
