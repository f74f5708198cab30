# developer tool (gpurun --gpus 8): weak scaling at 8 GPUs, rollout gather variants
run() { echo "$1: $(env $2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 192 --warmup 32 --no-cpu-baseline 2>gpurun_out/err_$3.log | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('value=%.4e ms_per_step=%.4f kernel_ms(min)=%.4f'%(j['value'], j['ms_per_step'], j['roofline']['kernel_ms']))")"; tail -2 gpurun_out/err_$3.log | grep -v "OMP\|\*\*\*" | cut -c1-300; }
mkdir -p gpurun_out
run "8 GPUs peer push (ring order, 4 copy streams), compact records" "UPKIE_BENCH_GATHER=peer" 29561
run "8 GPUs nccl, compact records" "UPKIE_BENCH_GATHER=nccl" 29562
