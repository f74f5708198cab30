# developer tool: time reduced-register builds of the step kernel (variants/ built by hand, see profiles/r01_variants.md)
mkdir -p gpurun_out
for spec in "n144 448" "r128 448" "n168 384" "n200 320"; do set -- $spec
UPKIE_B200_LIB=$PWD/variants/lib_$1.so UPKIE_B200_BLOCK=$2 timeout 150 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/regcap_$1.log 2>&1
echo "$1 block $2: $(tail -c 1500 gpurun_out/regcap_$1.log | grep -o '"value": [0-9.e+]*\|"kernel_ms": [0-9.]*\|Error.*\|error.*' | head -4 | tr '\n' ' ')"
done
