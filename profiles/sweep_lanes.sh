#!/bin/bash
# usage: profiles/sweep_lanes.sh out_file "L:C:over:conn" ...   (C3, 10 steps)
out=$1; shift
: > $out
for spec in "$@"; do
  IFS=: read L Cn OV CONN <<< "$spec"
  HP_DBG_OVERSUB=$OV CUDA_DEVICE_MAX_CONNECTIONS=${CONN:-32} python bench.py --steps 10 --warmup 3 --lanes $L --chunks $Cn --no-cpu-baseline ${EXTRA} 2>>$out.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$spec', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],1), 'alone', round(d['step_alone_ms'],1), 'frac', round(d['roofline']['frac'],3))
" >> $out
done
cat $out; tail -3 $out.err
