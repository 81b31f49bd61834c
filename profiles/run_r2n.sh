#!/bin/bash
# r2n: sub-solver key stripes of 6 instead of 8 AoS slots per lane in shared memory (6 KB instead of 8 KB per warp: a larger L1)
mkdir -p gpurun_out
: > gpurun_out/r2n_capl_s.txt
run() {  # label, env, args
  env $2 timeout 300 python bench.py $3 > gpurun_out/r2n_tmp.json 2> gpurun_out/r2n_tmp.err
  python - "$1" <<'PY' >> gpurun_out/r2n_capl_s.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/r2n_tmp.json").read().strip().splitlines()[-1])
    print("%-22s device-resident %.0f blocks/s (%.2f ms/step)  e2e %.0f (%.2f ms)  pageable %.0f  step alone %.1f ms" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_value") or 0, d.get("step_alone_ms", 0)))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run "c3 capl_s 6" "HP_B200_LIB=$PWD/build/libhp_s6.so" "--no-cpu-baseline"
run "c3 capl_s 8 (default)" "HP_X=1" "--no-cpu-baseline"
run "c3 capl_s 6 (2)" "HP_B200_LIB=$PWD/build/libhp_s6.so" "--no-cpu-baseline"
run "c2 capl_s 6" "HP_B200_LIB=$PWD/build/libhp_s6.so" "--config c2 --no-cpu-baseline"
run "c2 capl_s 8 (default)" "HP_X=1" "--config c2 --no-cpu-baseline"
cat gpurun_out/r2n_capl_s.txt
timeout 200 python profiles/ab.py default build/libhp_s6.so > gpurun_out/r2n_ab.txt 2>&1; cat gpurun_out/r2n_ab.txt
