#!/usr/bin/env python
"""C5-style run (BASELINE.json configs[4], scaled): the C3 block stream sharded over the GPUs of one box.
Every rank regenerates its contiguous shard from the seed (no input traffic), solves it through the C ABI, and the
per-block PhaseStats records are all-gathered over NCCL and re-ordered by block index (sharding.gather_block_records).
Timing: CUDA events per rank around the solve, max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      profiles/bench_c5.py [blocks_per_gpu]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from hiphase_b200 import lib, sharding, synth

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = per_gpu * world
lo, hi = sharding.contiguous_shard(total, rank, world)
t0 = time.time()
batch = synth.config_c3(hi - lo, first_block=lo)
gen_s = time.time() - t0
ctx = lib.Context(device=local)
ctx.astar_solve_batch(batch)                       # warm-up (allocations)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
out = ctx.astar_solve_batch(batch)                 # host buffers in / out: H2D + kernels + D2H
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
t = torch.tensor([wall, ctx.last_kernel_ms() / 1e3], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
rec = np.stack([out.stats[n].astype(np.int64) for n in out.stats.dtype.names], 1)
ids = np.arange(lo, hi, dtype=np.int64)
tg = time.perf_counter()
allrec = sharding.gather_block_records(ids, rec, total) if world > 1 else rec
gather_s = time.perf_counter() - tg
ok = int((out.status == 0).sum())
okt = torch.tensor([ok, batch.n_vars], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(okt, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"config": "C5 scaled: C3 block stream (N 20-2000, 30x, 2%% noisy), %d blocks per GPU" % per_gpu, "n_gpus": world,
                      "blocks": total, "variants": int(okt[1]), "blocks_ok": int(okt[0]), "e2e_s_max_over_ranks": float(t[0]),
                      "blocks_per_s": total / float(t[0]), "variants_per_s": int(okt[1]) / float(t[0]), "solver_kernels_s_max": float(t[1]),
                      "result_gather_s": gather_s, "gathered_records": int(allrec.shape[0]), "checksum_actual_cost": int(allrec[:, 2].sum()),
                      "shard_generation_s": gen_s}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
