#!/bin/bash
# Round-2 GPU call 10 (8 GPUs): the driver's scaling run - N = 1, 2, 4, 8 at --steps 20 --warmup 5 - on one box
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/scale_n1.json 2> $O/scale_n1.err
P=29700
for N in 2 4 8; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_n$N.json 2> $O/scale_n$N.err
done
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 256 --warmup 20 > $O/scale_n8_256.json 2> $O/scale_n8_256.err
timeout 300 python bench.py --gpus 1 --steps 256 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/scale_n1_256.json 2> $O/scale_n1_256.err
for f in n1 n2 n4 n8 n1_256 n8_256; do python - <<PY
import json
try:
    d=json.loads(open("$O/scale_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], "e2e %.3g"%d["e2e"]["value"], d["config"].get("gather"))
except Exception as e: print("$f failed", e)
PY
done
tail -n 3 $O/scale_n8.err
