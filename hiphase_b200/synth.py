"""Seeded synthetic workloads for the BASELINE.json configs.

Blocks are produced directly in the reference's u8 layout (clipped ReadSegments, src/data_types/read_segments.rs:40-62)
so the same batch feeds the CUDA path, the oracle and the pure-python restatement.
"""
import numpy as np

from ._abi import BlockBatch

# global-realignment quality table (src/read_parsing.rs:18-22, 815-835): SNV, indel, TR, SV
_TYPE_P = np.array([0.80, 0.15, 0.03, 0.02])
_TYPE_QUAL = np.array([160, 20, 80, 40], dtype=np.uint8)


def block_seed(config_id, block_id):
    return 0xB200 + config_id * 1000003 + block_id


def gen_block(rng, n_var, n_reads, span_fn, p_err, p_amb, p_gap=0.0, p_ignored=0.01, min_set=2):
    """One synthetic phase block -> dict(n_var, reads=[(start, alleles, quals)], ignored, is_snv)."""
    N = int(n_var)
    truth = rng.integers(0, 2, N, dtype=np.uint8)
    vtype = rng.choice(4, N, p=_TYPE_P)
    vqual = _TYPE_QUAL[vtype]
    is_snv = (vtype == 0).astype(np.uint8)
    ignored = (rng.random(N) < p_ignored)
    if N >= 2 and ignored.all():
        ignored[:] = False

    R = int(n_reads)
    spans = np.clip(span_fn(rng, R).astype(np.int64), 1, N)
    starts = np.floor(rng.random(R) * (N - spans + 1)).astype(np.int64)
    hap = rng.integers(0, 2, R, dtype=np.uint8)

    # make sure every variant sits inside at least one read (block_gen's connectivity): add short covering reads
    cover = np.zeros(N + 1, np.int64)
    np.add.at(cover, starts, 1)
    np.add.at(cover, starts + spans, -1)
    uncovered = np.flatnonzero(np.cumsum(cover[:N]) == 0)
    if len(uncovered) and N >= 3:
        es = np.clip(uncovered - 1, 0, N - 3)
        starts = np.concatenate([starts, es])
        spans = np.concatenate([spans, np.full(len(es), 3, np.int64)])
        hap = np.concatenate([hap, rng.integers(0, 2, len(es), dtype=np.uint8)])
        R = len(starts)

    total = int(spans.sum())
    first = np.concatenate([[0], np.cumsum(spans)[:-1]])
    rid = np.repeat(np.arange(R), spans)
    pos = starts[rid] + (np.arange(total) - first[rid])
    allele = truth[pos] ^ hap[rid]
    flip = rng.random(total) < p_err
    allele = np.where(flip, allele ^ 1, allele).astype(np.uint8)
    if p_amb > 0:
        allele[rng.random(total) < p_amb] = 2
    if p_gap > 0:
        allele[rng.random(total) < p_gap] = 3
    allele[ignored[pos]] = 3                     # astar_phaser.rs:435-442
    qual = np.where(allele < 2, vqual[pos], 0).astype(np.uint8)

    # ReadSegment::new clipping + --min-matched-alleles filter (read_parsing.rs:617)
    isset = allele < 2
    big = np.iinfo(np.int64).max
    pos_lo = np.where(isset, pos, big)
    pos_hi = np.where(isset, pos, -1)
    reads = []
    if total:
        lo = np.minimum.reduceat(pos_lo, first)
        hi = np.maximum.reduceat(pos_hi, first)
        nset = np.add.reduceat(isset.astype(np.int64), first)
        for r in np.flatnonzero(nset >= min_set):
            a0 = first[r] + (lo[r] - starts[r])
            a1 = first[r] + (hi[r] - starts[r]) + 1
            reads.append((int(lo[r]), allele[a0:a1], qual[a0:a1]))
    return {"n_var": N, "reads": reads, "ignored": ignored.astype(np.uint8), "is_snv": is_snv, "truth": truth}


def _uniform_span(lo, hi):
    return lambda rng, n: rng.integers(lo, hi + 1, n)


def _normal_span(mu, sd, lo, hi):
    return lambda rng, n: np.clip(np.rint(rng.normal(mu, sd, n)), lo, hi)


def config_c1(seed_base=1):
    """C1: single 50-variant x 30-read block (plumbing)."""
    rng = np.random.default_rng(block_seed(seed_base, 0))
    return BlockBatch.from_blocks([gen_block(rng, 50, 30, _uniform_span(8, 20), 0.01, 0.02)])


def c2_blocks(n_blocks=1000, first_block=0, n_var=200, n_reads=40, config_id=2):
    out = []
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(config_id, b))
        out.append(gen_block(rng, n_var, n_reads, _uniform_span(20, 40), 0.02, 0.03))
    return out


def config_c2(n_blocks=1000, first_block=0, n_var=200, n_reads=40):
    """C2: n_blocks independent blocks, 200 variants x 40 reads, spans U[20,40], p_err 0.02, p_amb 0.03."""
    return BlockBatch.from_blocks(c2_blocks(n_blocks, first_block, n_var, n_reads))


def config_c2_dense(n_blocks=1000, first_block=0):
    """C2 dense variant: 200 variants x 500 reads with spans ~12 (about 30x coverage)."""
    out = []
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(12, b))
        out.append(gen_block(rng, 200, 500, _normal_span(12, 4, 2, 40), 0.02, 0.03))
    return BlockBatch.from_blocks(out)


def c3_blocks(n_blocks=10000, first_block=0, n_lo=20, n_hi=2000, coverage=30, config_id=3):
    out = []
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(config_id, b))
        N = int(np.exp(rng.uniform(np.log(n_lo), np.log(n_hi))))
        noisy = rng.random() < 0.02
        R = max(2, int(coverage * N / 12.0))
        out.append(gen_block(rng, N, R, _normal_span(12, 4, 2, 40), 0.15 if noisy else 0.02, 0.03, 0.01))
    return out


def config_c3(n_blocks=10000, first_block=0, n_lo=20, n_hi=2000, coverage=30):
    """C3: HG002 chr20-scale: N log-uniform in [20,2000], 30x coverage, 2% noisy blocks."""
    return BlockBatch.from_blocks(c3_blocks(n_blocks, first_block, n_lo, n_hi, coverage))


def brute_force_mec(block):
    """Exhaustive minimum of sum_r min(score(h1), score(h2)) over all (h1,h2) in {0,1}^N x {0,1}^N (N <= 8)."""
    N = block["n_var"]
    assert N <= 8
    ign = block["ignored"].astype(bool)
    best = None
    # cost table per read per haplotype bit-pattern
    pats = np.arange(1 << N)
    bits = ((pats[:, None] >> np.arange(N)[None, :]) & 1).astype(np.uint8)   # [pat, var]
    tot = np.zeros((1 << N, 1 << N), np.int64)
    for (start, a, q) in block["reads"]:
        sc = np.zeros(1 << N, np.int64)
        for k in range(len(a)):
            i = start + k
            if ign[i]:
                continue
            sc += np.where(bits[:, i] != a[k], int(q[k]), 0)
        tot += np.minimum(sc[:, None], sc[None, :])
    best = int(tot.min())
    return best
