#!/usr/bin/env python
"""Graph-WFA realignment (BASELINE.json configs[3], "WFA-heavy") on one B200 next to the CPU oracle.
usage: python profiles/bench_wfa.py [n_blocks] [cpu_sample_jobs]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hiphase_b200 import lib, synth, _abi as A

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cpu_jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 96
t0 = time.time()
batch, jb, meta = synth.config_c4(nb)
print("generated %d jobs in %.1fs (%.1f MB reads)" % (batch.n_jobs, time.time() - t0, len(batch.read_bytes) / 1e6), flush=True)
ctx = lib.Context(device=0)
ctx.set_wfa_build_mode(int(os.environ.get("HP_WFA_BUILD_MODE", "0")))   # 0 = graphs built on the device, 1 = on the host
out = ctx.wfa_align_batch(batch, want_counters=True)
times = []
for _ in range(3):
    t0 = time.perf_counter(); out = ctx.wfa_align_batch(batch); times.append(time.perf_counter() - t0)
kms = ctx.last_kernel_ms()
e2e = min(times)
print("status counts", np.bincount(out.status, minlength=4).tolist())
res = {"jobs": batch.n_jobs, "e2e_s": e2e, "e2e_jobs_per_s": batch.n_jobs / e2e, "kernel_ms": kms,
       "kernel_jobs_per_s": batch.n_jobs / (kms / 1e3), "read_bytes": int(len(batch.read_bytes))}
import oracle_lib as O
sub_idx = np.arange(min(cpu_jobs, batch.n_jobs))
sub = A.WfaBatch({k: getattr(batch, k) for k in ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len",
                                                   "index_allele0", "vtype", "ignored", "allele_bytes")}, batch.reference,
                 batch.ref_start[sub_idx], batch.ref_end[sub_idx], batch.het_lo[sub_idx], batch.het_hi[sub_idx], batch.hom_lo[sub_idx],
                 batch.hom_hi[sub_idx], batch.read_bytes[: int(batch.read_off[len(sub_idx)])], batch.read_off[: len(sub_idx) + 1])
thr = os.cpu_count() or 1
t0 = time.perf_counter(); ref = O.wfa_align(sub, threads=thr, want_counters=True); dt = time.perf_counter() - t0
n = len(sub_idx)
ok = np.array_equal(ref.status, out.status[:n]) and np.array_equal(ref.score, out.score[:n]) and \
    np.array_equal(ref.alleles, out.alleles[: int(batch.row_off[n])]) and np.array_equal(ref.quals, out.quals[: int(batch.row_off[n])])
c = ref.counters
alg = int(2 * c["bases_compared"].sum() + 24 * c["waves_processed"].sum() + 8 * ((c["n_nodes"] + 63) // 64 * c["set_ops"]).sum())
res.update({"cpu_jobs": n, "cpu_s": dt, "cpu_jobs_per_s": n / dt, "cpu_threads": thr, "parity_on_sample": bool(ok),
            "speedup_e2e": (batch.n_jobs / e2e) / (n / dt), "alg_bytes_per_job_sample": alg / n})
print(json.dumps(res))
