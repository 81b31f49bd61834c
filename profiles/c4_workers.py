"""C4 pipeline with W worker threads (a context each): per-chunk wall times of the two calls, to see what overlaps.
usage: HP_DBG_REALIGN_TIMING=1 python profiles/c4_workers.py workers chunk"""
import os, sys, time, threading, queue
sys.path.insert(0, "/root/repo")
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
W, CH = int(sys.argv[1]), int(sys.argv[2])
chunks = [(A.RealignBatch(**d), vt) for (d, vt) in bench_c4.generate(0, 500, CH, os.cpu_count())]
arena = lib.PinnedArena()
for b, _ in chunks:
    for obj, names in ((b.wfa, ("read_bytes", "reference")), (b.local, ("read_bytes", "read_quals"))):
        for nme in names:
            setattr(obj, nme, arena.copy(getattr(obj, nme)))
snv = [np.concatenate([(np.array(v) == 0).astype(np.uint8) for v in vt]) for _, vt in chunks]
ctxs = [lib.Context(device=0) for _ in range(W)]
todo, log = queue.Queue(), []
def worker(w):
    c = ctxs[w]
    while True:
        k = todo.get()
        if k is None: return
        t0 = time.perf_counter(); r = c.realign_block_batch(chunks[k][0]); t1 = time.perf_counter()
        bb = r.block_batch(is_snv=snv[k]); t2 = time.perf_counter()
        c.astar_solve_batch(bb); t3 = time.perf_counter()
        log.append((w, k, t0, t1, t2, t3))
for rep in range(3):
    log.clear()
    th = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
    T0 = time.perf_counter()
    for k in range(len(chunks)): todo.put(k)
    for _ in th: todo.put(None)
    for t in th: t.start()
    for t in th: t.join()
    print("rep %d: step %.1f ms" % (rep, 1e3 * (time.perf_counter() - T0)), flush=True)
for w, k, t0, t1, t2, t3 in sorted(log, key=lambda x: x[2]):
    print("worker %d chunk %d: realign [%6.1f, %6.1f] glue %.1f astar [%6.1f, %6.1f] ms" % (w, k, 1e3 * (t0 - T0), 1e3 * (t1 - T0), 1e3 * (t2 - t1), 1e3 * (t2 - T0), 1e3 * (t3 - T0)))
