#!/bin/bash
# Round 2, session 2 (2 GPUs): the in-kernel rollout transport with the final kernels: transport check + bench at the driver's settings
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/multicast_check.py > $O/multicast_check.log 2>&1; echo "rc=$?" >> $O/multicast_check.log
grep -h "rollout\|rc=\|MATCH" $O/multicast_check.log | sort | uniq -c | head
timeout 300 $TR --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
timeout 300 $TR --master-port 29591 bench.py --gpus 2 --steps 256 --warmup 20 --no-cpu-baseline > $O/bench_n2_256.json 2> $O/bench_n2_256.err
timeout 200 $TR --master-port 29601 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref rc=$?"
for f in bench_n2 bench_n2_256 bench_ref_n2; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d.get("ms_per_step",0), "kernel", (d.get("roofline") or {}).get("kernel_ms"), (d["config"].get("gather") or {}))
except Exception as e: print("$f failed", e)
PY
done
tail -n 3 $O/bench_n2.err
