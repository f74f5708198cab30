#!/bin/bash
# Round-2 GPU call 2 (2 GPUs): in-kernel rollout transports (multicast / peer stores) validated and timed.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi topo -m > $O/topo2.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/multicast_check.py > $O/multicast_check2.log 2>&1; echo "rc=$?" >> $O/multicast_check2.log
timeout 200 python bench.py --steps 256 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/bench2_n1.json 2> $O/bench2_n1.err
P=29600
for G in multicast peerstore peer nccl; do
  P=$((P+1))
  UPKIE_BENCH_GATHER=$G timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 256 --warmup 20 > $O/bench2_$G.json 2> $O/bench2_$G.err
done
timeout 300 $TR --master-port 29620 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench2_driver.json 2> $O/bench2_driver.err
cat $O/multicast_check2.log | tail -8
for f in n1 multicast peerstore peer nccl driver; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench2_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], "e2e %.3g"%d["e2e"]["value"], d["config"].get("gather"), d["clocks"])
except Exception as e: print("$f failed", e)
PY
done
tail -5 $O/bench2_multicast.err $O/bench2_peerstore.err
