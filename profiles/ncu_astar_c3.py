#!/usr/bin/env python
"""Workload for the ncu capture of astar_solve_kernel on the C3 stream: n blocks in one batch, team 1 (the streaming shape).
usage: ncu ... python profiles/ncu_astar_c3.py [n_blocks] [team]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
batch = synth.config_c3_stream(nb)
ctx = lib.Context(device=0)
ctx.set_team(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
for _ in range(2):
    out = ctx.astar_solve_batch(batch, want_counters=(_ == 0))
    if _ == 0:
        print("pops", int(out.counters["pops"].sum()), "evals", int(out.counters["evals"].sum()), "cells", int(out.counters["cells"].sum()))
print("kernel ms", ctx.last_kernel_ms(), "status ok", int((out.status == 0).sum()))
