#!/usr/bin/env python
"""A* throughput on the other BASELINE.json configs (C3 subset, C2 dense) on one B200, with oracle parity on a sample.
usage: python profiles/bench_configs.py [c3_blocks]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hiphase_b200 import lib, synth
import oracle_lib as O

def run(name, batch, sample):
    ctx = lib.Context(device=0)
    ctx.astar_solve_batch(batch)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out = ctx.astar_solve_batch(batch); ts.append(time.perf_counter() - t0)
    kms = ctx.last_kernel_ms()
    idx = np.linspace(0, batch.n_blocks - 1, sample).astype(int)
    sub = batch.select(idx)
    thr = os.cpu_count() or 1
    t0 = time.perf_counter(); ref = O.astar_solve(sub, threads=thr, want_heuristic=False, want_counters=False); dt = time.perf_counter() - t0
    ok = all(np.array_equal(out.h1[int(batch.var_off[i]):int(batch.var_off[i + 1])], ref.h1[int(sub.var_off[k]):int(sub.var_off[k + 1])]) and out.stats[i] == ref.stats[k]
             for k, i in enumerate(idx))
    print(json.dumps({"config": name, "blocks": batch.n_blocks, "variants": batch.n_vars, "cells": batch.n_cells, "e2e_blocks_per_s": batch.n_blocks / min(ts),
                      "e2e_variants_per_s": batch.n_vars / min(ts), "solver_kernels_ms": kms, "status_ok": int((out.status == 0).sum()),
                      "pruned_blocks": int((out.stats["pruned_solutions"] > 0).sum()),
                      "cpu_sample_blocks": sample, "cpu_threads": thr, "cpu_blocks_per_s": sample / dt,
                      "cpu_variants_per_s": sub.n_vars / dt, "parity_on_sample": bool(ok)}), flush=True)
    ctx.close()

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
t0 = time.time(); b3 = synth.config_c3(n3); print("generated C3 subset in %.1fs" % (time.time() - t0), flush=True)
run("C3 subset (N log-uniform 20-2000, 30x coverage, 2%% noisy): %d blocks" % n3, b3, 48)
run("C2 dense (200 variants x 500 reads): 200 blocks", synth.config_c2_dense(200), 24)
