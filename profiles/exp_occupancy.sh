run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/r2d_exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'alone', round(d['step_alone_ms'],1))
"; }
run base A=1
run ctas5 HP_B200_LIB=build/alt/libhp_ctas5.so
run w12 HP_DBG_WARPS_PER_SM=12
run w8 HP_DBG_WARPS_PER_SM=8
tail -3 gpurun_out/r2d_exp.err
