#!/bin/bash
# r2k: GPU tests (multi-panel edit distance, full-size C3 test), 20 warps per SM (build/libhp_c5.so: 96 registers) vs the default
# 16, C4 / C2 bench lines with the final library.  Run under gpurun from the repo root.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2k_pytest.log
timeout 300 python profiles/ab.py default build/libhp_c5.so > gpurun_out/r2k_ab.txt 2>&1; cat gpurun_out/r2k_ab.txt
HP_B200_LIB=$PWD/build/libhp_c5.so timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench_c3_c5.json 2> gpurun_out/r2k_bench_c3_c5.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err
timeout 400 python bench.py --config c4 > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err
timeout 200 python bench.py --config c2 > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err
python - <<'PY'
import json
for f in ("r2k_bench_c3_c5", "r2k_bench_c3", "r2k_bench_c4", "r2k_bench_c2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f pageable %.0f ms/step %.2f alone %.1f parity %s" % (d["value"], d["e2e"]["value"], d["e2e"].get("pageable_value") or 0, d["ms_per_step"], d.get("step_alone_ms", 0), d.get("parity_sample")))
    except Exception as e:
        print(f, "failed", e)
PY
