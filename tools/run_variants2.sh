# developer tool: barrier / block-size variants of the device-buffer step kernel
run() { echo "$1: $(env $2 timeout 150 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('value=%.4e kernel_ms=%.4f'%(j['value'], j['roofline']['kernel_ms']))")"; }
run "default (barrier, block 224)" "A=1"
run "barrier, block 128" "UPKIE_B200_BLOCK=128"
run "barrier, block 160" "UPKIE_B200_BLOCK=160"
run "barrier, block 256" "UPKIE_B200_BLOCK=256"
run "barrier, block 64" "UPKIE_B200_BLOCK=64"
run "no barrier, block 224" "UPKIE_B200_LIB=$PWD/variants/lib_nosync.so"
run "no barrier, block 128" "UPKIE_B200_LIB=$PWD/variants/lib_nosync.so UPKIE_B200_BLOCK=128"
run "no barrier, block 64" "UPKIE_B200_LIB=$PWD/variants/lib_nosync.so UPKIE_B200_BLOCK=64"
run "no barrier, block 32" "UPKIE_B200_LIB=$PWD/variants/lib_nosync.so UPKIE_B200_BLOCK=32"
