import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
chunks = [(A.RealignBatch(**d), vt) for (d, vt) in bench_c4.generate(0, 500, 250, os.cpu_count())]
ctx = lib.Context(device=0)
arena = lib.PinnedArena()
for b, _ in chunks:
    for obj, names in ((b.wfa, ("read_bytes", "reference")), (b.local, ("read_bytes", "read_quals"))):
        for nme in names:
            setattr(obj, nme, arena.copy(getattr(obj, nme)))
for rep in range(3):
    t = {"realign": 0, "glue": 0, "astar": 0}
    for b, vt in chunks:
        t0 = time.perf_counter(); r = ctx.realign_block_batch(b); t1 = time.perf_counter()
        is_snv = np.concatenate([(np.array(v) == 0).astype(np.uint8) for v in vt]); bb = r.block_batch(is_snv=is_snv); t2 = time.perf_counter()
        o = ctx.astar_solve_batch(bb); t3 = time.perf_counter()
        t["realign"] += t1 - t0; t["glue"] += t2 - t1; t["astar"] += t3 - t2
    print({k: round(v * 1e3, 1) for k, v in t.items()}, "astar kernel ms", ctx.last_kernel_ms())
