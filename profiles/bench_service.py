#!/usr/bin/env python
"""The unchanged-orchestrator shape: T worker threads, each calling hp_service_solve_one for one phase block at a time
(src/main.rs:385-408), on the C3 stream.  usage: python profiles/bench_service.py [n_blocks] [threads,...]"""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
tlist = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 64, 256]
batch = synth.stream_blocks(np.arange(0, 10000, 10000 // nb, dtype=np.uint64)[:nb])
for T in tlist:
    svc = lib.BlockService(device=0)
    nxt = [0]
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                b = nxt[0]; nxt[0] += 1
            if b >= nb:
                return
            svc.solve_one(batch, b)
    for warm in (True, False):
        nxt[0] = 0
        th = [threading.Thread(target=worker) for _ in range(T)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
    nbat, nblk = svc.counters()
    print(json.dumps({"threads": T, "blocks": nb, "blocks_per_s": nb / dt, "batches": nbat, "blocks_per_batch": nblk / max(nbat, 1)}), flush=True)
    svc.close()
