#!/bin/bash
# Round-2 GPU call 12b (2 GPUs): in-kernel transports re-validated (transport exact, physics to round-off); push variants timed
mkdir -p gpurun_out/r02
O=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/multicast_check.py > $O/multicast_check2c.log 2>&1; echo "rc=$?" >> $O/multicast_check2c.log
grep -h "rollout\|rc=" $O/multicast_check2c.log | sort
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/b12_n1.json 2> $O/b12_n1.err
P=29800
for V in kernel now deferred; do
  P=$((P+1))
  UPKIE_BENCH_PUSH=$V timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $O/b12_$V.json 2> $O/b12_$V.err
done
P=$((P+1))
UPKIE_BENCH_PUSH=kernel timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 256 --warmup 20 > $O/b12_kernel_256.json 2> $O/b12_kernel_256.err
for f in n1 kernel now deferred kernel_256; do python - <<PY
import json
try:
    d=json.loads(open("$O/b12_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], (d["config"].get("gather") or {}).get("transport"))
except Exception as e: print("$f failed", e)
PY
done
tail -n 4 $O/b12_kernel.err
