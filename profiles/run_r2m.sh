#!/bin/bash
# r2m: the dense (20 warps per SM) build of the production kernels chosen per launch in the throughput regime.
# GPU tests, C3 bench with / without it (HP_DBG_DENSE=0), C2 line, ncu captures of both builds of astar_solve_kernel<2>.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2m_pytest.log
: > gpurun_out/r2m_dense_ab.txt
run() {  # label, env
  env $2 timeout 300 python bench.py $3 > gpurun_out/r2m_tmp.json 2> gpurun_out/r2m_tmp.err
  python - "$1" <<'PY' >> gpurun_out/r2m_dense_ab.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/r2m_tmp.json").read().strip().splitlines()[-1])
    print("%-22s device-resident %.0f blocks/s (%.2f ms/step)  e2e %.0f (%.2f ms)  pageable %.0f  step alone %.1f ms  parity %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_value") or 0, d.get("step_alone_ms", 0), d.get("parity_sample")))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run "c3 dense (default)" "HP_X=1" ""; cp gpurun_out/r2m_tmp.json gpurun_out/r2m_bench_c3.json
run "c3 HP_DBG_DENSE=0" "HP_DBG_DENSE=0" "--no-cpu-baseline"
run "c3 dense (default) 2" "HP_X=1" "--no-cpu-baseline"
run "c3 HP_DBG_DENSE=0 2" "HP_DBG_DENSE=0" "--no-cpu-baseline"
run "c2 dense (default)" "HP_X=1" "--config c2"; cp gpurun_out/r2m_tmp.json gpurun_out/r2m_bench_c2.json
run "c2 HP_DBG_DENSE=0" "HP_DBG_DENSE=0" "--config c2 --no-cpu-baseline"
cat gpurun_out/r2m_dense_ab.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:astar_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2m_astar_c3 python profiles/ncu_astar_c3.py 10000 1 > gpurun_out/r2m_ncu_astar.log 2>&1
HP_DBG_DENSE=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:astar_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2m_astar_c3_dense python profiles/ncu_astar_c3.py 10000 1 > gpurun_out/r2m_ncu_astar_dense.log 2>&1
tail -2 gpurun_out/r2m_ncu_astar.log gpurun_out/r2m_ncu_astar_dense.log
