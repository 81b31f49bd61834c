#!/bin/bash
# r2q: final GPU tests of the round + device share (HP_DBG_OVERSUB) and lanes with the dense build, C3 stream
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2q_pytest.log
: > gpurun_out/r2q_share_lanes.txt
run() {  # label, env, args
  env $2 timeout 300 python bench.py $3 > gpurun_out/r2q_tmp.json 2> gpurun_out/r2q_tmp.err
  python - "$1" <<'PY' >> gpurun_out/r2q_share_lanes.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/r2q_tmp.json").read().strip().splitlines()[-1])
    print("%-26s device-resident %.0f blocks/s (%.2f ms/step)  e2e %.0f (%.2f ms)  pageable %.0f  step alone %.1f ms" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_value") or 0, d.get("step_alone_ms", 0)))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run "over 1.5 lanes 8 (default)" "HP_X=1" ""; cp gpurun_out/r2q_tmp.json gpurun_out/r2q_bench_c3.json
run "over 1.25 lanes 8" "HP_DBG_OVERSUB=1.25" "--no-cpu-baseline"
run "over 2.0 lanes 8" "HP_DBG_OVERSUB=2.0" "--no-cpu-baseline"
run "over 1.5 lanes 12" "HP_X=1" "--no-cpu-baseline --lanes 12"
run "over 1.5 lanes 6" "HP_X=1" "--no-cpu-baseline --lanes 6"
cat gpurun_out/r2q_share_lanes.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_c3_reference.json 2>/dev/null; tail -c 400 gpurun_out/r2q_bench_c3_reference.json
