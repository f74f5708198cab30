# developer tool (gpurun --gpus 2): rollout gather implementation / record size vs weak-scaling efficiency
run() { echo "$1: $(env $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 320 --warmup 32 --no-cpu-baseline 2>gpurun_out/err_$3.log | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('value=%.4e ms_per_step=%.4f kernel_ms(min)=%.4f'%(j['value'], j['ms_per_step'], j['roofline']['kernel_ms']))")"; tail -2 gpurun_out/err_$3.log | grep -v "OMP\|\*\*\*" | cut -c1-300; }
mkdir -p gpurun_out
timeout 200 python bench.py --steps 320 --warmup 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('1 GPU compact: value=%.4e ms_per_step=%.4f'%(j['value'], j['ms_per_step']))"
UPKIE_BENCH_ROLLOUT=full timeout 200 python bench.py --steps 320 --warmup 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('1 GPU full: value=%.4e ms_per_step=%.4f'%(j['value'], j['ms_per_step']))"
run "peer push, compact records" "UPKIE_BENCH_GATHER=peer" 29541
run "peer push, full records" "UPKIE_BENCH_GATHER=peer UPKIE_BENCH_ROLLOUT=full" 29542
run "nccl, compact records" "UPKIE_BENCH_GATHER=nccl" 29543
run "peer push, compact again" "UPKIE_BENCH_GATHER=peer" 29544
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
