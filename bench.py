#!/usr/bin/env python
"""bench.py -- phase blocks/sec of the per-block phasing hot path on B200 (BASELINE.json metric).

A "step" is one pass of the A* hot path (prep + solver kernels) over one batch of synthetic phase blocks.
Workload at N GPUs: BASELINE.json configs[1] per rank ("1k independent blocks, 200 variants x 40 reads"), i.e.
weak scaling over independent blocks with no data-path collective (blocks share nothing, src/main.rs:385-408).

  value      blocks/s with the batch already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        blocks/s through the C-ABI host entry (hp_astar_solve_batch): pinned host buffers, H2D + kernels + D2H
  roofline   algorithmic bytes of the solver kernel (SURVEY.md 8d formula, counted exactly by the kernel itself and
             checked against the oracle's counters in tests/) / its CUDA-event duration, vs the measured HBM peak
  cpu_baseline  the CPU oracle (reference-equivalent C++ restatement; the Rust reference cannot be built in this
             image) on a bounded sample of the same workload, all host cores, one block per worker task

  --impl reference   times that CPU restatement as the reference arm (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_BLOCKS = 1000          # BASELINE.json configs[1]
N_VAR, N_READS = 200, 40
CPU_SAMPLE_BLOCKS = 384


def algorithmic_bytes(counters):
    """per child evaluation sum_r 2*w_r*3 B + (4*L_parent + 2) B + 40 B."""
    return int(6 * counters["cells"].sum() + 4 * counters["sum_parent_len"].sum() + 42 * counters["evals"].sum())


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_oracle_blocks_per_s(first_block, n_blocks, threads):
    """The CPU restatement (oracle) on n_blocks blocks of the same workload.  Checker / baseline only."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from hiphase_b200 import synth
    batch = synth.config_c2(n_blocks=n_blocks, first_block=first_block, n_var=N_VAR, n_reads=N_READS)
    t0 = time.perf_counter()
    out = O.astar_solve(batch, threads=threads, want_heuristic=False, want_counters=False)
    dt = time.perf_counter() - t0
    assert out.failures == 0
    return n_blocks / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cpu_oracle_blocks_per_s(0, 32, threads)
    for _ in range(max(0, args.warmup - 1)):
        cpu_oracle_blocks_per_s(0, 32, threads)
    times = []
    for k in range(args.steps):
        v, dt = cpu_oracle_blocks_per_s(k * CPU_SAMPLE_BLOCKS, CPU_SAMPLE_BLOCKS, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = CPU_SAMPLE_BLOCKS / (ms / 1e3)
    sample = "%d blocks of the same workload per step (bounded sample), one block per worker task" % CPU_SAMPLE_BLOCKS
    line = {"impl": "reference", "metric": "phase blocks/sec", "value": value, "unit": "blocks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic",
            "config": {"workload": "C2: %d-variant x %d-read phase blocks (BASELINE.json configs[1])" % (N_VAR, N_READS),
                       "note": "CPU restatement of astar_phaser.rs (oracle/); the Rust reference cannot be built here"},
            "cpu_baseline": {"value": value, "unit": "blocks/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=N_BLOCKS, help="phase blocks per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from hiphase_b200 import _abi as A
    from hiphase_b200 import lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hiphase_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- this rank's shard of the workload (independent blocks: no input traffic between ranks) ----
    nb = args.blocks
    batch = synth.config_c2(n_blocks=nb, first_block=rank * nb, n_var=N_VAR, n_reads=N_READS)
    max_n = int(np.diff(batch.var_off.astype(np.int64)).max())
    ctx = lib.Context(device=local_rank)
    if os.environ.get("HP_TEAM"):
        ctx.set_team(int(os.environ["HP_TEAM"]))      # experiment knob: speculative team size (default: automatic)

    # device-resident copy of the batch (reference u8 layout) + device outputs
    dten = {k: torch.from_numpy(getattr(batch, k).view(np.int64) if getattr(batch, k).dtype == np.uint64 else
                                (getattr(batch, k).view(np.int32) if getattr(batch, k).dtype == np.uint32 else getattr(batch, k))).to(dev)
            for k in A.BlockBatch.FIELDS}

    def dptr(t, ty):
        return C.cast(t.data_ptr(), ty)
    dbatch = A.hp_block_batch(nb, dptr(dten["var_off"], A.u64p), dptr(dten["read_off"], A.u64p), dptr(dten["read_start"], A.u32p),
                              dptr(dten["read_end"], A.u32p), dptr(dten["cell_off"], A.u64p), dptr(dten["alleles"], A.u8p),
                              dptr(dten["quals"], A.u8p), dptr(dten["ignored"], A.u8p), dptr(dten["is_snv"], A.u8p))
    o_h1 = torch.empty(batch.n_vars, dtype=torch.uint8, device=dev)
    o_h2 = torch.empty(batch.n_vars, dtype=torch.uint8, device=dev)
    o_stats = torch.zeros(nb * 7, dtype=torch.int64, device=dev)
    o_status = torch.full((nb,), -1, dtype=torch.int32, device=dev)
    o_ctr = torch.zeros(nb * 4, dtype=torch.int64, device=dev)
    dout_ctr = A.hp_astar_out(dptr(o_h1, A.u8p), dptr(o_h2, A.u8p), C.cast(o_stats.data_ptr(), C.POINTER(A.hp_phase_stats)),
                              dptr(o_status, A.i32p), A.u64p(), C.cast(o_ctr.data_ptr(), C.POINTER(A.hp_astar_counters)))
    # the timed steps run the production kernel (no work counters); one warm-up step runs the counting variant
    dout = A.hp_astar_out(dptr(o_h1, A.u8p), dptr(o_h2, A.u8p), C.cast(o_stats.data_ptr(), C.POINTER(A.hp_phase_stats)),
                          dptr(o_status, A.i32p), A.u64p(), C.POINTER(A.hp_astar_counters)())
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device(o=None):
        ctx.astar_solve_device(dbatch, batch.n_vars, batch.n_reads, batch.n_cells, max_n, o or dout, torch.cuda.current_stream().cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    step_device(dout_ctr)
    for _ in range(args.warmup):
        flush.fill_(1)
        step_device()
    sync_all()
    assert int((o_status != 0).sum().item()) == 0, "blocks failed on the device"
    ctr = o_ctr.cpu().numpy().view(np.uint64).reshape(nb, 4)
    counters = {"evals": ctr[:, 0], "cells": ctr[:, 1], "sum_parent_len": ctr[:, 2], "pops": ctr[:, 3]}
    alg_bytes = algorithmic_bytes(counters)

    # ---- timed region: K steps, device-resident inputs, L2 flushed between steps (flush excluded via events) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    sync_all()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        step_device()
        ev[k][1].record()
        kernel_ms.append(None)
        ev[k][1].synchronize()
        kernel_ms[k] = ctx.last_kernel_ms()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = nb * world / (ms_per_step / 1e3)

    # ---- end to end through the host C-ABI entry: pinned host inputs, H2D + kernels + D2H in the timed region ----
    pinned = {k: torch.from_numpy(getattr(batch, k).view(np.uint8)).pin_memory() for k in A.BlockBatch.FIELDS}
    hb = A.BlockBatch.__new__(A.BlockBatch)
    for k in A.BlockBatch.FIELDS:
        setattr(hb, k, pinned[k].numpy().view(getattr(batch, k).dtype))
    hb.n_blocks = nb
    h2d = sum(int(pinned[k].numel()) for k in A.BlockBatch.FIELDS)
    d2h = 2 * batch.n_vars + nb * (56 + 4)
    for _ in range(2):
        ctx.astar_solve_batch(hb)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eo = ctx.astar_solve_batch(hb)
    e2e_s = (time.perf_counter() - t0) / args.steps
    assert (eo.status == 0).all()
    assert np.array_equal(eo.h1, o_h1.cpu().numpy()) and np.array_equal(eo.h2, o_h2.cpu().numpy())
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = nb * world / e2e_s
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- result hand-off between ranks (outside the timed region): the only collective on the path.  Per-block
    #      PhaseStats records are all-gathered over NCCL and re-ordered by global block index.
    checksum = int(o_h1.to(torch.int64).sum().item() * 3 + o_h2.to(torch.int64).sum().item())
    if world > 1:
        from hiphase_b200 import sharding
        ids = np.arange(rank * nb, (rank + 1) * nb, dtype=np.int64)
        stats_all = sharding.gather_block_records(ids, o_stats.cpu().numpy().reshape(nb, 7), nb * world)
        assert stats_all.shape == (nb * world, 7) and (stats_all[:, 2] >= stats_all[:, 1]).all()
        t = torch.tensor([checksum], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        checksum = int(t.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k_ms = float(np.mean([m for m in kernel_ms if m]))
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (profiles/)
        traffic = None
        try:
            tb = 0.0
            import glob
            # the latest committed capture of this kernel (profiles/r1*_ncu_astar_solve_kernel.txt sort by name)
            for ln in open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r1*_ncu_astar_solve_kernel.txt")))[-1]):
                f = ln.split()
                if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tb += float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            traffic = tb or None
        except Exception:
            pass
        achieved = alg_bytes / (k_ms / 1e3) / 1e9
        line = {"metric": "phase blocks/sec", "value": value, "unit": "blocks/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/u32 (integer)", "data": "synthetic",
                "config": {"workload": "C2: %d independent phase blocks per GPU, %d variants x %d reads (BASELINE.json configs[1])" % (nb, N_VAR, N_READS),
                           "blocks_per_gpu": nb, "cells_per_gpu": batch.n_cells, "parallelism": "blocks sharded over %d GPU(s), no data-path collective" % world,
                           "l2": "192 MiB flush buffer written between timed steps (inputs are smaller than L2)",
                           "params": {"min_queue_size": 1000, "queue_increment": 3}},
                "e2e": {"value": e2e_value, "unit": "blocks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s, "api": "hp_astar_solve_batch (host buffers, pinned)"},
                "gpu_launches": int(launches), "wall_s_timed_region": t_wall,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "astar_solve_kernel", "kernel_ms": k_ms,
                             "algorithmic_bytes_per_launch": alg_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                             "note": "algorithmic bytes = the reference's u8 rescoring traffic (6 B/cell + node clones), counted by the kernel and equal to the oracle's counters; the working set is L1/L2/SMEM resident so DRAM traffic is far lower by design"},
                "clocks": sampler.summary(), "result_checksum": checksum}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_oracle_blocks_per_s(0, 16, threads)
            v, dt = cpu_oracle_blocks_per_s(0, CPU_SAMPLE_BLOCKS, threads)
            line["cpu_baseline"] = {"value": v, "unit": "blocks/s", "cores": threads, "kind": "port",
                                    "sample": "first %d blocks of the same workload, %.1f s wall, one block per worker task" % (CPU_SAMPLE_BLOCKS, dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
