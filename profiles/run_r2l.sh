#!/bin/bash
# r2l: resident warps per SM in the streaming regime: 16 (default) vs 20 (build/libhp_c5.so) vs 24 (libhp_c6.so), alternating runs
mkdir -p gpurun_out
: > gpurun_out/r2l_warps_per_sm.txt
for rep in 1 2 3; do
  for v in default c5 c6; do
    if [ $v = c6 ] && [ $rep != 1 ]; then continue; fi
    if [ $v = default ]; then unset HP_B200_LIB; else export HP_B200_LIB=$PWD/build/libhp_$v.so; fi
    timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2l_tmp.json 2> gpurun_out/r2l_tmp.err
    python - $v $rep <<'PY' >> gpurun_out/r2l_warps_per_sm.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/r2l_tmp.json").read().strip().splitlines()[-1])
    print("%-8s rep %s  device-resident %.0f blocks/s (%.2f ms/step)  e2e %.0f (%.2f ms)  pageable %.0f  step alone %.1f ms" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_value") or 0, d.get("step_alone_ms", 0)))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e)
PY
  done
done
cat gpurun_out/r2l_warps_per_sm.txt
