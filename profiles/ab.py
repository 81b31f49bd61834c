#!/usr/bin/env python
"""A/B timing of kernel variants (same ABI, different builds): python profiles/ab.py lib1.so lib2.so ...
Each variant runs in its own process: C2 (1000 blocks) and a C3 subset, solver-kernel ms (CUDA events), median of 7."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from hiphase_b200 import lib, synth
    ctx = lib.Context(device=0)
    if os.environ.get("AB_TEAM"):
        ctx.set_team(int(os.environ["AB_TEAM"]))
    res = []
    for name, batch in (("c2", synth.config_c2(1000)), ("c3", synth.config_c3(int(os.environ.get("AB_C3", "1000"))))):
        ms = []
        for _ in range(8):
            out = ctx.astar_solve_batch(batch)
            ms.append(ctx.last_kernel_ms())
        res.append("%s %.3f ms (min %.3f) chk %d" % (name, float(np.median(ms[1:])), min(ms[1:]), int(out.stats["actual_cost"].sum() % 1000003)))
    print(os.environ.get("HP_B200_LIB", "default"), " | ".join(res), flush=True)
else:
    for so in sys.argv[1:]:
        env = dict(os.environ)
        if so != "default":
            env["HP_B200_LIB"] = os.path.abspath(so)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env)
