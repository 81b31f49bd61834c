"""Graph-WFA kernel time split: jobs that align (ED ~ 30) vs jobs that run into MaxEditDistance (5 % error reads)."""
import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from hiphase_b200 import lib, synth
ctx = lib.Context(device=0)
ctx.set_wfa_filter(os.environ.get("HP_WFA_FILTER", "1") != "0")
for name, kw in (("normal only", dict(p_noisy=0.0)), ("noisy only", dict(p_noisy=1.0)), ("mixed", dict())):
    nb = 64 if name != "noisy only" else 8
    batch, jb, meta = synth.config_c4(nb, **kw)
    ctx.wfa_align_batch(batch)
    ctx.wfa_align_batch(batch)
    print(name, "production kernel ms %.2f, filter answered %d" % (ctx.last_kernel_ms(), ctx.wfa_filtered()))
    t0 = time.perf_counter(); out = ctx.wfa_align_batch(batch, want_counters=True); dt = time.perf_counter() - t0
    c = out.counters
    print(name, "jobs", batch.n_jobs, "kernel ms %.2f" % ctx.last_kernel_ms(), "e2e ms %.1f" % (dt * 1e3), "status", np.bincount(out.status, minlength=2).tolist(),
          "mean score %.1f waves %.0f cmp %.0f" % (out.score.mean(), c["waves_processed"].mean(), c["bases_compared"].mean()), flush=True)
