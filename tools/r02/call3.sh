#!/bin/bash
# Round-2 GPU call 3 (8 GPUs): in-kernel rollout transports at 8 ranks, against NCCL and the copy-engine push.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi topo -m > $O/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 tools/multicast_check.py > $O/multicast_check8.log 2>&1; echo "rc=$?" >> $O/multicast_check8.log
timeout 200 python bench.py --steps 256 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/bench8_n1.json 2> $O/bench8_n1.err
P=29600
for G in multicast peerstore nccl peer; do
  P=$((P+1))
  UPKIE_BENCH_GATHER=$G timeout 300 $TR --master-port $P bench.py --gpus 8 --steps 256 --warmup 20 > $O/bench8_$G.json 2> $O/bench8_$G.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/bench8_n1_driver.json 2> $O/bench8_n1_driver.err
timeout 300 $TR --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench8_driver.json 2> $O/bench8_driver.err
grep -h "rollout\|rc=" $O/multicast_check8.log | sort | uniq -c | head -8
for f in n1 multicast peerstore nccl peer n1_driver driver; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench8_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"], "e2e %.3g"%d["e2e"]["value"], d["config"].get("gather"))
except Exception as e: print("$f failed", e)
PY
done
for f in multicast peerstore; do tail -n 3 $O/bench8_$f.err; done
