"""A* stage of the C4 pipeline: what the 250 blocks of a chunk cost (kernel ms, wall, per-block cycles, retries)."""
import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 250
(d, vt), = bench_c4.generate(0, nb, nb, os.cpu_count())
b = A.RealignBatch(**d)
ctx = lib.Context(device=0)
r = ctx.realign_block_batch(b)
is_snv = np.concatenate([(np.array(v) == 0).astype(np.uint8) for v in vt])
bb = r.block_batch(is_snv=is_snv)
nvar = np.diff(bb.var_off.astype(np.int64)); nrd = np.diff(bb.read_off.astype(np.int64))
print("blocks %d, variants/block %.1f, reads/block %.1f, cells %d" % (bb.n_blocks, nvar.mean(), nrd.mean(), int(bb.cell_off[-1])))
L = lib.lib()
for rep in range(3):
    l0 = ctx.launch_count(); t0 = time.perf_counter(); o = ctx.astar_solve_batch(bb); dt = time.perf_counter() - t0
    print("astar_solve_batch wall %.1f ms, last kernel %.2f ms, launches %d" % (dt * 1e3, ctx.last_kernel_ms(), ctx.launch_count() - l0))
L.hp_debug_enable_block_cycles(ctx.handle, 1)
o = ctx.astar_solve_batch(bb, want_counters=True)
dd = np.zeros(nb * 16, np.uint64)
assert L.hp_debug_read_block_cycles(ctx.handle, dd.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
dd = dd.reshape(nb, 16).astype(np.float64)
tot = dd[:, 0] + dd[:, 1]
print("counting variant kernel %.2f ms; per block cycles: mean %.3g median %.3g max %.3g; pre-pass share %.2f" % (ctx.last_kernel_ms(), tot.mean(), np.median(tot), tot.max(), dd[:, 0].sum() / tot.sum()))
print("stats pruned>0 blocks:", int((o.stats["pruned_solutions"] > 0).sum()) if o.stats.dtype.names else "n/a")
for k in np.argsort(tot)[-5:]:
    print("  block %d: cycles %.3g (pre %.3g main %.3g) pops pre %d main %d, N %d R %d" % (k, tot[k], dd[k, 0], dd[k, 1], dd[k, 2], dd[k, 3], nvar[k], nrd[k]))
