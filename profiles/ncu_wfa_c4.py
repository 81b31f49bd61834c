#!/usr/bin/env python
"""Workload for the ncu capture of wfa_align_kernel: n blocks of the C4 pipeline workload, graph-WFA jobs only.
usage: ncu ... python profiles/ncu_wfa_c4.py [n_blocks]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 50
d, vt = bench_c4.generate(0, nb, nb, os.cpu_count() or 1)[0]
ctx = lib.Context(device=0)
for k in range(2):
    out = ctx.wfa_align_batch(d["wfa"], want_counters=(k == 0))
    if k == 0:
        c = out.counters
        alg = int(2 * c["bases_compared"].sum() + 24 * c["waves_processed"].sum() + 8 * ((c["n_nodes"] + 63) // 64 * c["set_ops"]).sum())
        ok = out.status == 0
        print("jobs", d["wfa"].n_jobs, "status", np.bincount(out.status, minlength=5).tolist(), "alg bytes", alg,
              "mean score ok %.1f" % out.score[ok].mean(), "waves/job ok %.0f" % c["waves_processed"][ok].mean(), "fail %.0f" % c["waves_processed"][~ok].mean(),
              "cmp/job ok %.0f" % c["bases_compared"][ok].mean(), "nodes mean %.0f" % c["n_nodes"].mean())
print("kernel ms", ctx.last_kernel_ms())
