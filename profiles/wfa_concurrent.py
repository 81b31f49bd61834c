"""Graph-WFA launches of W contexts side by side (one host thread each) on chunks of 250 C4 blocks: wall time per round."""
import os, sys, time, threading
sys.path.insert(0, "/root/repo")
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
chunks = [A.RealignBatch(**d) for (d, vt) in bench_c4.generate(0, 500, 250, os.cpu_count())]
arena = lib.PinnedArena()
for b in chunks:
    for nme in ("read_bytes", "reference"):
        setattr(b.wfa, nme, arena.copy(getattr(b.wfa, nme)))
for W in (1, 2, 4):
    ctxs = [lib.Context(device=0) for _ in range(W)]
    outs = [A.WfaOut(chunks[w % 2].wfa) for w in range(W)]
    def work(w, reps):
        for _ in range(reps):
            ctxs[w].wfa_align_batch(chunks[w % 2].wfa)
    for w in range(W): work(w, 1)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(w, 4)) for w in range(W)]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print("W=%d: %d calls of 37500 jobs in %.1f ms -> %.1f ms per call (last kernel %.1f ms)" % (W, 4 * W, 1e3 * dt, 1e3 * dt / (4 * W), ctxs[0].last_kernel_ms()), flush=True)
    for c in ctxs: c.close()
