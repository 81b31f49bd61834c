#!/bin/bash
# r2j: packed-u16x2 sub-solver vectors (default build) vs the unpacked build (build/libhp_unpacked.so), GPU tests, C3 bench,
# refreshed ncu capture of the solver kernel.  Run under gpurun from the repo root.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2j_pytest.log
timeout 300 python profiles/ab.py default build/libhp_unpacked.so > gpurun_out/r2j_ab.txt 2>&1; cat gpurun_out/r2j_ab.txt
timeout 400 python bench.py > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; echo "bench exit $?"
HP_B200_LIB=$PWD/build/libhp_unpacked.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_c3_unpacked.json 2> gpurun_out/r2j_bench_c3_unpacked.err
python - <<'PY'
import json
for f in ("r2j_bench_c3", "r2j_bench_c3_unpacked"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f pageable %.0f ms/step %.2f alone %.1f parity %s" % (d["value"], d["e2e"]["value"], d["e2e"].get("pageable_value", 0), d["ms_per_step"], d.get("step_alone_ms", 0), d.get("parity_sample")))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:astar_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2j_astar_c3 python profiles/ncu_astar_c3.py 10000 1 > gpurun_out/r2j_ncu_astar.log 2>&1
tail -2 gpurun_out/r2j_ncu_astar.log
ls -la gpurun_out | tail -12
