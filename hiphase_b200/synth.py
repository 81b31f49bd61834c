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


# ---------------------------------------------------------------------------------------------------------------
# HG002-scale stream (C3 = its first 10 000 blocks, C5 = its first 200 000): multi-threaded C++ generator
# (csrc/hp_synth.cpp -> libhp_synth.so).  Block b depends on (config_id, b) alone, so ranks regenerate their own shards.
# ---------------------------------------------------------------------------------------------------------------
import ctypes as _C
import os as _os

_SYNTH_LIB = None


class hp_synth_params(_C.Structure):
    _fields_ = [("config_id", _C.c_uint64), ("n_lo", _C.c_uint32), ("n_hi", _C.c_uint32), ("coverage", _C.c_double),
                ("p_noisy", _C.c_double), ("p_err", _C.c_double), ("p_err_noisy", _C.c_double), ("p_amb", _C.c_double),
                ("p_gap", _C.c_double), ("p_ignored", _C.c_double)]


def _synth_lib():
    global _SYNTH_LIB
    if _SYNTH_LIB is None:
        from . import _abi as A
        path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "csrc", "libhp_synth.so")
        L = _C.CDLL(path)
        L.hp_synth_default_params.argtypes = [_C.POINTER(hp_synth_params)]
        L.hp_synth_headers.argtypes = [_C.POINTER(hp_synth_params), _C.c_uint64, _C.c_uint64, A.u32p, A.u8p]
        L.hp_synth_generate.argtypes = [_C.POINTER(hp_synth_params), A.u64p, _C.c_uint64, _C.c_int, _C.POINTER(_C.c_void_p)]
        L.hp_synth_view.argtypes = [_C.c_void_p, _C.POINTER(A.hp_block_batch)]
        L.hp_synth_free.argtypes = [_C.c_void_p]
        _SYNTH_LIB = L
    return _SYNTH_LIB


def stream_params(**kw):
    p = hp_synth_params()
    _synth_lib().hp_synth_default_params(_C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def stream_headers(first_block, n_blocks, params=None):
    """(n_var[n], noisy[n]) of blocks [first_block, first_block + n) of the stream, without generating them."""
    from . import _abi as A
    p = params or stream_params()
    nv = np.zeros(n_blocks, np.uint32)
    noisy = np.zeros(n_blocks, np.uint8)
    _synth_lib().hp_synth_headers(_C.byref(p), int(first_block), int(n_blocks), A.ptr(nv, A.u32p), A.ptr(noisy, A.u8p))
    return nv, noisy


def stream_blocks(ids, params=None, threads=None, alloc=None):
    """The stream's blocks `ids` (any order) as a BlockBatch.  alloc(nbytes) -> writable uint8 array (e.g. pinned memory)."""
    from . import _abi as A
    p = params or stream_params()
    ids = np.ascontiguousarray(ids, np.uint64)
    h = _C.c_void_p()
    L = _synth_lib()
    L.hp_synth_generate(_C.byref(p), A.ptr(ids, A.u64p), len(ids), int(threads or _os.cpu_count() or 1), _C.byref(h))
    try:
        v = A.hp_block_batch()
        L.hp_synth_view(h, _C.byref(v))
        nb = len(ids)

        def take(ptr_, n, dt):
            dt = np.dtype(dt)
            if n == 0:
                return np.zeros(0, dt)
            src = np.ctypeslib.as_array(ptr_, shape=(n,))
            if alloc is None:
                return src.astype(dt, copy=True)
            dst = alloc(n * dt.itemsize).view(dt)
            dst[:] = src
            return dst
        var_off = take(v.var_off, nb + 1, np.uint64); read_off = take(v.read_off, nb + 1, np.uint64)
        nv, nr = int(var_off[-1]), int(read_off[-1])
        cell_off = take(v.cell_off, nr + 1, np.uint64)
        nc = int(cell_off[-1])
        b = BlockBatch.__new__(BlockBatch)
        b.var_off, b.read_off, b.cell_off = var_off, read_off, cell_off
        b.read_start = take(v.read_start, nr, np.uint32); b.read_end = take(v.read_end, nr, np.uint32)
        b.alleles = take(v.alleles, nc, np.uint8); b.quals = take(v.quals, nc, np.uint8)
        b.ignored = take(v.ignored, nv, np.uint8); b.is_snv = take(v.is_snv, nv, np.uint8)
        b.n_blocks = nb
        return b
    finally:
        L.hp_synth_free(h)


def config_c3_stream(n_blocks=10000, first_block=0, threads=None):
    """C3 / C5: blocks [first_block, first_block + n_blocks) of the HG002-scale stream."""
    return stream_blocks(np.arange(first_block, first_block + n_blocks, dtype=np.uint64), threads=threads)


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


# ---------------------------------------------------------------------------------------------------------------
# C4: WFA-heavy workload (graph realignment jobs), SURVEY.md section 8d
# ---------------------------------------------------------------------------------------------------------------
_VT_SNV, _VT_INS, _VT_DEL, _VT_INDEL, _VT_SVINS, _VT_SVDEL, _VT_TR = 0, 1, 2, 3, 4, 5, 9
_BASES = np.frombuffer(b"ACGT", np.uint8)


def _rand_seq(rng, n):
    return _BASES[rng.integers(0, 4, n)]


def gen_wfa_block(rng, window=75000, n_het=60, n_hom=90, n_reads=150, read_lo=10000, read_hi=20000,
                  err=0.002, p_noisy=0.03, err_noisy=0.05, sv_max=2000):
    """One block: reference window, het/hom variants (truncated alleles), reads sampled from the two haplotypes.
    Returns dict(reference, hets=[...], homs=[...], jobs=[(ref_start, ref_end, het_lo, het_hi, hom_lo, hom_hi, read)])."""
    ref = _rand_seq(rng, window)
    nv = n_het + n_hom
    pos = np.sort(rng.choice(np.arange(50, window - sv_max - 300), nv, replace=False))
    is_het = np.zeros(nv, bool)
    is_het[rng.choice(nv, n_het, replace=False)] = True
    variants = []
    for i in range(nv):
        p = int(pos[i])
        u = rng.random()
        if u < 0.5:
            rl = 1
            alt = _BASES[(int(np.searchsorted(_BASES, ref[p])) + int(rng.integers(1, 4))) % 4: ][:1]
            v = dict(vtype=_VT_SNV, ref_len=1, a0=ref[p:p + 1], a1=np.array(alt, np.uint8), i0=0)
        elif u < 0.8:
            kind = int(rng.integers(0, 3))
            k = int(rng.integers(1, 21))
            if kind == 0:
                v = dict(vtype=_VT_INS, ref_len=1, a0=ref[p:p + 1], a1=np.concatenate([ref[p:p + 1], _rand_seq(rng, k)]), i0=0)
            elif kind == 1:
                v = dict(vtype=_VT_DEL, ref_len=k + 1, a0=ref[p:p + k + 1], a1=ref[p:p + 1], i0=0)
            else:
                r = int(rng.integers(2, 11))
                v = dict(vtype=_VT_INDEL, ref_len=r, a0=ref[p:p + r], a1=np.concatenate([ref[p:p + 1], _rand_seq(rng, int(rng.integers(1, 10)))]), i0=0)
        elif u < 0.85:      # multi-allelic indel: both alleles are ALTs (index_allele0 = 1, index_allele1 = 2)
            r = int(rng.integers(2, 8))
            v = dict(vtype=_VT_INDEL, ref_len=r, a0=np.concatenate([ref[p:p + 1], _rand_seq(rng, int(rng.integers(1, 6)))]),
                     a1=np.concatenate([ref[p:p + 1], _rand_seq(rng, int(rng.integers(6, 12)))]), i0=1)
        elif u < 0.95:      # tandem-repeat like expansion of a short motif
            motif = _rand_seq(rng, int(rng.integers(2, 7)))
            L = int(rng.integers(10, 60))
            v = dict(vtype=_VT_TR, ref_len=L, a0=ref[p:p + L],
                     a1=np.concatenate([ref[p:p + L], np.tile(motif, int(rng.integers(2, 30)))[: int(rng.integers(10, 201))]]), i0=0)
        else:
            L = int(rng.integers(50, sv_max + 1))
            if rng.random() < 0.5:
                v = dict(vtype=_VT_SVINS, ref_len=1, a0=ref[p:p + 1], a1=np.concatenate([ref[p:p + 1], _rand_seq(rng, L)]), i0=0)
            else:
                v = dict(vtype=_VT_SVDEL, ref_len=L + 1, a0=ref[p:p + L + 1], a1=ref[p:p + 1], i0=0)
        v["pos"] = p
        v["het"] = bool(is_het[i])
        v["phase"] = int(rng.integers(0, 2))
        variants.append(v)
    hets = [v for v in variants if v["het"]]
    homs = [v for v in variants if not v["het"]]

    # the two haplotype sequences + reference->haplotype coordinate maps at backbone positions
    haps, maps = [], []
    for h in (0, 1):
        out, cur, mp_ref, mp_hap, hp_len = [], 0, [0], [0], 0
        for v in variants:
            if v["pos"] < cur:
                continue                      # overlaps the variant applied just before on this haplotype
            allele = v["a1"] if not v["het"] else (v["a1"] if v["phase"] == h else v["a0"])
            seg = ref[cur:v["pos"]]
            out.append(seg); hp_len += len(seg)
            mp_ref.append(v["pos"]); mp_hap.append(hp_len)
            out.append(allele); hp_len += len(allele)
            cur = v["pos"] + v["ref_len"]
            mp_ref.append(cur); mp_hap.append(hp_len)
        out.append(ref[cur:]); hp_len += window - cur
        mp_ref.append(window); mp_hap.append(hp_len)
        haps.append(np.concatenate(out))
        maps.append((np.array(mp_ref), np.array(mp_hap)))

    het_pos = np.array([v["pos"] for v in hets])
    hom_pos = np.array([v["pos"] for v in homs])
    spans = np.array([(v["pos"], v["pos"] + v["ref_len"]) for v in variants])

    def to_hap(h, x):   # x must be a backbone position (outside every applied variant span)
        mr, mh = maps[h]
        k = int(np.searchsorted(mr, x, side="right")) - 1
        return int(mh[k] + (x - mr[k]))

    def free_pos(x):    # move x right until it is not inside any variant's reference span
        for _ in range(64):
            inside = (spans[:, 0] <= x) & (x < spans[:, 1] + 1)
            if not inside.any():
                return x
            x = int(spans[inside, 1].max()) + 1
        return x

    jobs = []
    for _ in range(n_reads):
        h = int(rng.integers(0, 2))
        ln = int(rng.integers(read_lo, read_hi + 1))
        s = free_pos(int(rng.integers(0, max(1, window - ln))))
        e = free_pos(min(window - 1, s + ln))
        if e >= window or e <= s + 10:
            continue
        seq = haps[h][to_hap(h, s): to_hap(h, e)].copy()
        er = err_noisy if rng.random() < p_noisy else err
        n = len(seq)
        # substitutions, deletions, insertions at rate er/3 each
        sub = rng.random(n) < er / 3
        seq[sub] = _BASES[rng.integers(0, 4, int(sub.sum()))]
        keep = rng.random(n) >= er / 3
        ins = rng.random(n) < er / 3
        pieces = np.where(keep, 1, 0) + np.where(ins, 1, 0)
        outseq = np.empty(int(pieces.sum()), np.uint8)
        idx = np.cumsum(pieces) - pieces
        outseq[idx[keep]] = seq[keep]
        ins_at = idx[ins] + np.where(keep[ins], 1, 0)
        outseq[ins_at] = _BASES[rng.integers(0, 4, int(ins.sum()))]
        jobs.append((s, e, int(np.searchsorted(het_pos, s)), int(np.searchsorted(het_pos, e)),
                     int(np.searchsorted(hom_pos, s)), int(np.searchsorted(hom_pos, e)), outseq, h))
    return dict(reference=ref, hets=hets, homs=homs, jobs=jobs)


def config_c4(n_blocks=500, first_block=0, **kw):
    """C4: WFA-heavy realignment jobs.  Returns (WfaBatch, job_block[ n_jobs ], blocks_meta)."""
    from ._abi import WfaBatch
    refs, ref_base = [], 0
    vt = {k: [] for k in ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len", "index_allele0", "vtype", "ignored")}
    blob = []
    blob_len = 0
    rs, re, hl, hh, ml, mh, reads, job_block, meta = [], [], [], [], [], [], [], [], []
    nvar = 0
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(4, b))
        blk = gen_wfa_block(rng, **kw)
        het_base = nvar
        for v in blk["hets"] + blk["homs"]:
            vt["position"].append(v["pos"] + ref_base); vt["ref_len"].append(v["ref_len"])
            vt["allele0_off"].append(blob_len); vt["allele0_len"].append(len(v["a0"])); blob.append(v["a0"]); blob_len += len(v["a0"])
            vt["allele1_off"].append(blob_len); vt["allele1_len"].append(len(v["a1"])); blob.append(v["a1"]); blob_len += len(v["a1"])
            vt["index_allele0"].append(v["i0"]); vt["vtype"].append(v["vtype"]); vt["ignored"].append(0)
            nvar += 1
        hom_base = het_base + len(blk["hets"])
        for (s, e, a, bb, c, d, seq, h) in blk["jobs"]:
            rs.append(s + ref_base); re.append(e + ref_base)
            hl.append(het_base + a); hh.append(het_base + bb); ml.append(hom_base + c); mh.append(hom_base + d)
            reads.append(seq); job_block.append(b - first_block)
        meta.append(dict(n_het=len(blk["hets"]), het_base=het_base, vtypes=[v["vtype"] for v in blk["hets"]],
                         phase=[v["phase"] for v in blk["hets"]]))
        refs.append(blk["reference"]); ref_base += len(blk["reference"])
    vt["allele_bytes"] = np.concatenate(blob) if blob else np.zeros(0, np.uint8)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])])
    batch = WfaBatch(vt, np.concatenate(refs), rs, re, hl, hh, ml, mh, np.concatenate(reads), read_off)
    return batch, np.array(job_block), meta


def blocks_from_wfa_rows(batch, out, job_block, meta, min_set=2):
    """WFA rows -> A* phase blocks (one ReadSegment per job; reads with < min_set set alleles are dropped,
    read_parsing.rs:612-629).  Jobs that hit MaxEditDistance (status 1) or were skipped contribute nothing."""
    blocks = [dict(n_var=m["n_het"], reads=[], is_snv=(np.array(m["vtypes"]) == 0).astype(np.uint8)) for m in meta]
    for j in range(batch.n_jobs):
        if out.status[j] != 0:
            continue
        m = meta[job_block[j]]
        r0, r1 = int(batch.row_off[j]), int(batch.row_off[j + 1])
        a, q = out.alleles[r0:r1], out.quals[r0:r1]
        sset = np.flatnonzero(a < 2)
        if len(sset) < min_set:
            continue
        lo, hi = int(sset[0]), int(sset[-1]) + 1
        start = int(batch.het_lo[j]) - m["het_base"] + lo
        blocks[job_block[j]]["reads"].append((start, a[lo:hi].copy(), q[lo:hi].copy()))
    return BlockBatch.from_blocks(blocks)


# ---- local realignment jobs ---------------------------------------------------------------------
def extend_with_reference(variants, ref, reference_buffer=15):
    """Reference prefix / postfix of the het variants exactly as load_variant_calls does it (src/phaser.rs:236-296):
    up to `reference_buffer` bases either side, the previous variant's postfix truncated where two variants crowd."""
    previous_het_end = 0
    for i, v in enumerate(variants):
        pos, ref_len = v.position(), v.get_ref_len()
        ref_prefix_start = pos - reference_buffer if pos > reference_buffer else 0
        ref_postfix_start = pos + ref_len
        if ref_prefix_start < previous_het_end:
            prev = variants[i - 1]
            current_end = prev.position() + prev.get_ref_len() + prev.get_postfix_len()
            prev.truncate_reference_postfix(min(max(current_end - pos, 0), prev.get_postfix_len()))
            ref_prefix_start = min(previous_het_end, pos)
        v.add_reference_prefix(ref[ref_prefix_start:pos].tobytes())
        v.add_reference_postfix(ref[ref_postfix_start:ref_postfix_start + reference_buffer].tobytes())
        previous_het_end = pos + ref_len


def gen_local_block(rng, window=30000, n_var=40, n_reads=60, read_lo=4000, read_hi=15000, err=0.003, sv_max=1500,
                    reference_buffer=15, p_ignored=0.02, n_hom=0, p_noisy=0.0, err_noisy=0.05):
    """One block for local realignment: reference window, het variants (with reference prefix / postfix), reads sampled
    from the two haplotypes and their alignments (gap-free aligned segments, the M/=/X runs of a CIGAR).
    n_hom extra homozygous variants are carried by every read (they only matter to the graph of global realignment);
    a fraction p_noisy of the reads is drawn with error rate err_noisy.
    Returns dict(variants=[Variant] (hets), homs=[Variant], jobs=[(var_lo, var_hi, read_pos, segs[(ref, read, len)], seq, quals)])."""
    from .variants import Variant
    ref = _rand_seq(rng, window)
    n_all = n_var + n_hom
    pos = np.sort(rng.choice(np.arange(100, window - sv_max - 300), n_all, replace=False))
    is_hom = np.zeros(n_all, bool)
    if n_hom:
        is_hom[rng.choice(n_all, n_hom, replace=False)] = True
    allv, phase = [], []
    for p in (int(x) for x in pos):
        u = rng.random()
        if u < 0.6:
            alt = _BASES[(int(np.searchsorted(_BASES, ref[p])) + int(rng.integers(1, 4))) % 4]
            v = Variant(0, _VT_SNV, p, 1, ref[p:p + 1].tobytes(), bytes([alt]))
        elif u < 0.8:
            kind, k = int(rng.integers(0, 3)), int(rng.integers(1, 21))
            if kind == 0:
                v = Variant(0, _VT_INS, p, 1, ref[p:p + 1].tobytes(), ref[p:p + 1].tobytes() + _rand_seq(rng, k).tobytes())
            elif kind == 1:
                v = Variant(0, _VT_DEL, p, k + 1, ref[p:p + k + 1].tobytes(), ref[p:p + 1].tobytes())
            else:
                r = int(rng.integers(2, 11))
                v = Variant(0, _VT_INDEL, p, r, ref[p:p + r].tobytes(), ref[p:p + 1].tobytes() + _rand_seq(rng, int(rng.integers(1, 10))).tobytes())
        elif u < 0.85:     # multi-allelic indel, both alleles ALT
            r = int(rng.integers(2, 8))
            v = Variant(0, _VT_INDEL, p, r, ref[p:p + 1].tobytes() + _rand_seq(rng, int(rng.integers(1, 6))).tobytes(),
                        ref[p:p + 1].tobytes() + _rand_seq(rng, int(rng.integers(6, 12))).tobytes(), 1, 2)
        elif u < 0.92:
            motif, L = _rand_seq(rng, int(rng.integers(2, 7))), int(rng.integers(10, 60))
            exp = np.tile(motif, 40)[: int(rng.integers(10, 201))]
            v = Variant(0, _VT_TR, p, L, ref[p:p + L].tobytes(), ref[p:p + L].tobytes() + exp.tobytes())
        else:
            L = int(rng.integers(50, sv_max + 1))
            if rng.random() < 0.5:
                v = Variant(0, _VT_SVINS, p, 1, ref[p:p + 1].tobytes(), ref[p:p + 1].tobytes() + _rand_seq(rng, L).tobytes())
            else:
                v = Variant(0, _VT_SVDEL, p, L + 1, ref[p:p + L + 1].tobytes(), ref[p:p + 1].tobytes())
        allv.append(v); phase.append(int(rng.integers(0, 2)))
    raw = [(v.get_allele0(), v.get_allele1()) for v in allv]      # alleles before the reference extension
    vs = [v for v, h in zip(allv, is_hom) if not h]
    homs = [v for v, h in zip(allv, is_hom) if h]
    extend_with_reference(vs, ref, reference_buffer)
    for v in vs:
        if rng.random() < p_ignored:
            v.set_ignored()
    vpos = np.array([v.position() for v in vs])

    jobs = []
    for _ in range(n_reads):
        h = int(rng.integers(0, 2))
        e_rate = err_noisy if (p_noisy > 0 and rng.random() < p_noisy) else err
        ln = int(rng.integers(read_lo, read_hi + 1))
        s = int(rng.integers(0, max(1, window - ln)))
        e = min(window, s + ln)
        seq, segs = [], []              # seq: list of uint8 chunks
        rd = 0                          # read cursor
        run = None                      # open aligned run [ref_start, read_start, len]

        def close():
            nonlocal run
            if run is not None and run[2] > 0:
                segs.append(tuple(run))
            run = None

        def add_run(ref_pos, cnt):
            nonlocal rd, run
            if cnt <= 0:
                return
            if run is None:
                run = [ref_pos, rd, 0]
            run[2] += cnt
            seq.append(ref[ref_pos:ref_pos + cnt]); rd += cnt

        def emit_match(lo, hi):
            nonlocal rd
            n = hi - lo
            if n <= 0:
                return
            r = rng.random(n)
            start = 0
            for ev in np.flatnonzero(r < 2 * e_rate / 3):
                ev = int(ev)
                add_run(lo + start, ev - start)
                close()
                if r[ev] < e_rate / 3:                # deletion error: reference base without a read base
                    start = ev + 1
                else:                                 # insertion error before this base
                    seq.append(_BASES[rng.integers(0, 4, 1)]); rd += 1
                    start = ev
            add_run(lo + start, n - start)

        cur = s
        for k, v in enumerate(allv):
            p, rl = v.position(), v.get_ref_len()
            if p < cur or p + rl > e:
                continue
            emit_match(cur, p)
            al = np.frombuffer(raw[k][1] if (is_hom[k] or phase[k] == h) else raw[k][0], np.uint8)
            m = min(len(al), rl)
            if run is None:
                run = [p, rd, 0]
            run[2] += m
            seq.append(al[:m]); rd += m
            if len(al) > m:                          # insertion
                close()
                seq.append(al[m:]); rd += len(al) - m
            elif rl > m:                             # deletion
                close()
            cur = p + rl
        emit_match(cur, e)
        close()
        if not segs:
            continue
        seq = np.concatenate(seq).astype(np.uint8)
        sub = rng.random(len(seq)) < e_rate / 3
        seq[sub] = _BASES[rng.integers(0, 4, int(sub.sum()))]
        q = rng.integers(15, 60, len(seq)).astype(np.uint8)
        low = rng.random(len(seq)) < 0.03
        q[low] = rng.integers(0, 10, int(low.sum()))
        jobs.append((int(np.searchsorted(vpos, s)), int(np.searchsorted(vpos, e)), segs[0][0], segs, seq, q))
    return dict(variants=vs, homs=homs, jobs=jobs, reference=ref)


def config_local(n_blocks=8, first_block=0, full_rows=False, **kw):
    """Local-realignment batch over n_blocks blocks.  full_rows=True gives every job the block's whole variant list (the
    reference calls local_realignment with all variants of the block, read_parsing.rs:568)."""
    from ._abi import LocalBatch
    from .variants import variant_table
    allv, var_lo, var_hi, read_pos, seg_off, sr, sd, sl, reads, quals = [], [], [], [], [0], [], [], [], [], []
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(6, b))
        blk = gen_local_block(rng, **kw)
        base = len(allv)
        allv.extend(blk["variants"])
        for (lo, hi, rp, segs, seq, q) in blk["jobs"]:
            var_lo.append(base if full_rows else base + lo)
            var_hi.append(base + len(blk["variants"]) if full_rows else base + hi)
            read_pos.append(rp)
            for (a, r, n) in segs:
                sr.append(a); sd.append(r); sl.append(n)
            seg_off.append(len(sr))
            reads.append(seq); quals.append(q)
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])])
    return LocalBatch(variant_table(allv), var_lo, var_hi, read_pos, seg_off, sr, sd, sl,
                      np.concatenate(reads), np.concatenate(quals), read_off)


# ---- realignment pipeline batches: the same mappings as graph-WFA jobs AND local-realignment jobs ---------------------
def config_realign(n_blocks=4, first_block=0, p_pair=0.15, **kw):
    """Blocks of gen_local_block with, per mapping, its global-realignment job (plan_global_realignment: window, overlapped
    hets, aligned read slice; truncated alleles) and its local-realignment job (all hets of the block; full alleles).
    A fraction p_pair of the mappings shares its read name with the previous mapping (supplementary alignments -> collapse).
    Returns (RealignBatch kwargs dict, vtypes per block)."""
    from ._abi import LocalBatch, WfaBatch
    from .read_parsing import AlignedRead, plan_global_realignment
    from .variants import variant_table
    allv = []
    wt = {k: [] for k in ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len", "index_allele0", "vtype", "ignored")}
    blob, blob_len = [], 0
    refs, ref_base = [], 0
    var_lo, var_hi, read_pos, seg_off, sr, sd, sl, l_reads, l_quals = [], [], [], [0], [], [], [], [], []
    rs, re, hl, hh, ml, mh, w_reads = [], [], [], [], [], [], []
    map_off, map_group, n_groups, var_off, het_base, vtypes = [0], [], [], [0], [], []
    for b in range(first_block, first_block + n_blocks):
        rng = np.random.default_rng(block_seed(8, b))
        blk = gen_local_block(rng, p_ignored=0.0, **kw)
        vs = blk["variants"]
        base = len(allv)
        allv.extend(vs)
        wbase = len(wt["position"])                                 # the WFA table interleaves per block: hets, then homs
        het_base.append(wbase)
        for v in vs:
            a0, a1 = v.get_truncated_allele0(), v.get_truncated_allele1()
            wt["position"].append(v.position() + ref_base); wt["ref_len"].append(v.get_ref_len())
            wt["allele0_off"].append(blob_len); wt["allele0_len"].append(len(a0)); blob.append(np.frombuffer(a0, np.uint8)); blob_len += len(a0)
            wt["allele1_off"].append(blob_len); wt["allele1_len"].append(len(a1)); blob.append(np.frombuffer(a1, np.uint8)); blob_len += len(a1)
            wt["index_allele0"].append(v.index_allele0); wt["vtype"].append(int(v.get_type())); wt["ignored"].append(0)
        vpos = [v.position() for v in vs]
        homs = blk.get("homs", [])
        hom_first = len(wt["position"])
        for v in homs:                                               # hom calls follow the block's hets in the WFA table
            a0, a1 = v.get_allele0(), v.get_allele1()
            wt["position"].append(v.position() + ref_base); wt["ref_len"].append(v.get_ref_len())
            wt["allele0_off"].append(blob_len); wt["allele0_len"].append(len(a0)); blob.append(np.frombuffer(a0, np.uint8)); blob_len += len(a0)
            wt["allele1_off"].append(blob_len); wt["allele1_len"].append(len(a1)); blob.append(np.frombuffer(a1, np.uint8)); blob_len += len(a1)
            wt["index_allele0"].append(v.index_allele0); wt["vtype"].append(int(v.get_type())); wt["ignored"].append(0)
        hpos = [v.position() for v in homs]
        g = -1
        for (lo, hi, rp, segs, seq, q) in blk["jobs"]:
            if g < 0 or rng.random() >= p_pair:
                g += 1
            map_group.append(g)
            var_lo.append(base); var_hi.append(base + len(vs)); read_pos.append(rp)
            for (a, r, n) in segs:
                sr.append(a); sd.append(r); sl.append(n)
            seg_off.append(len(sr))
            l_reads.append(seq); l_quals.append(q)
            plan = plan_global_realignment(AlignedRead(rp, segs, seq.tobytes(), q), vpos, hpos)
            if plan is None:
                rs.append(ref_base); re.append(ref_base); hl.append(wbase); hh.append(wbase); ml.append(hom_first); mh.append(hom_first)
                w_reads.append(np.zeros(0, np.uint8))
            else:
                rs.append(plan["ref_start"] + ref_base); re.append(plan["ref_end"] + ref_base)
                hl.append(wbase + plan["het_lo"]); hh.append(wbase + plan["het_hi"])
                ml.append(hom_first + plan["hom_lo"]); mh.append(hom_first + plan["hom_hi"])
                w_reads.append(seq[plan["read_start"]:plan["read_end"]])
        map_off.append(len(map_group)); n_groups.append(g + 1); var_off.append(var_off[-1] + len(vs))
        vtypes.append([int(v.get_type()) for v in vs])
        refs.append(blk["reference"]); ref_base += len(blk["reference"])
    wt["allele_bytes"] = np.concatenate(blob) if blob else np.zeros(1, np.uint8)
    nm = len(map_group)
    w_off = np.concatenate([[0], np.cumsum([len(r) for r in w_reads])])
    wfa = WfaBatch(wt, np.concatenate(refs), rs, re, hl, hh, ml, mh,
                   np.concatenate(w_reads) if w_off[-1] else np.zeros(1, np.uint8), w_off)
    l_off = np.concatenate([[0], np.cumsum([len(r) for r in l_reads])])
    local = LocalBatch(variant_table(allv), var_lo, var_hi, read_pos, seg_off, sr, sd, sl, np.concatenate(l_reads), np.concatenate(l_quals), l_off)
    return dict(wfa=wfa, local=local, map_off=map_off, map_group=map_group, n_groups=n_groups, var_off=var_off, wfa_het_base=het_base), vtypes


def merge_realign(pieces):
    """Concatenates config_realign results (generated piecewise, e.g. by a process pool) into one: [(kwargs, vtypes)] -> same."""
    from ._abi import LocalBatch, WfaBatch
    if len(pieces) == 1:
        return pieces[0]
    cat = np.concatenate
    vkeys = ("ref_len", "allele0_len", "allele1_len", "index_allele0", "vtype", "ignored")
    wt = {k: [] for k in vkeys + ("position", "allele0_off", "allele1_off", "allele_bytes")}
    lt = {k: [] for k in vkeys + ("position", "allele0_off", "allele1_off", "allele_bytes", "prefix_len", "postfix_len")}
    w = {k: [] for k in ("reference", "ref_start", "ref_end", "het_lo", "het_hi", "hom_lo", "hom_hi", "read_bytes", "read_off")}
    l = {k: [] for k in ("var_lo", "var_hi", "read_pos", "seg_off", "seg_ref_start", "seg_read_start", "seg_len", "read_bytes", "read_quals", "read_off")}
    map_off, map_group, n_groups, var_off, het_base, vtypes = [np.zeros(1, np.uint64)], [], [], [np.zeros(1, np.uint64)], [], []
    o = dict(ref=0, wab=0, wnv=0, wrb=0, lab=0, lnv=0, lseg=0, lrb=0, nm=0, nvar=0)
    for d, vt in pieces:
        W, L = d["wfa"], d["local"]
        for k in vkeys:
            wt[k].append(getattr(W, k)); lt[k].append(getattr(L, k))
        wt["position"].append(W.position + o["ref"]); lt["position"].append(L.position)
        wt["allele0_off"].append(W.allele0_off + np.uint64(o["wab"])); wt["allele1_off"].append(W.allele1_off + np.uint64(o["wab"]))
        lt["allele0_off"].append(L.allele0_off + np.uint64(o["lab"])); lt["allele1_off"].append(L.allele1_off + np.uint64(o["lab"]))
        wt["allele_bytes"].append(W.allele_bytes); lt["allele_bytes"].append(L.allele_bytes)
        lt["prefix_len"].append(L.prefix_len); lt["postfix_len"].append(L.postfix_len)
        w["reference"].append(W.reference)
        w["ref_start"].append(W.ref_start + np.uint64(o["ref"])); w["ref_end"].append(W.ref_end + np.uint64(o["ref"]))
        for k in ("het_lo", "het_hi", "hom_lo", "hom_hi"):
            w[k].append(getattr(W, k) + np.uint32(o["wnv"]))
        w["read_bytes"].append(W.read_bytes[: int(W.read_off[-1])]); w["read_off"].append(W.read_off[:-1] + np.uint64(o["wrb"]))
        l["var_lo"].append(L.var_lo + np.uint32(o["lnv"])); l["var_hi"].append(L.var_hi + np.uint32(o["lnv"]))
        l["read_pos"].append(L.read_pos)
        l["seg_off"].append(L.seg_off[:-1] + np.uint64(o["lseg"]))
        for k in ("seg_ref_start", "seg_read_start", "seg_len"):
            l[k].append(getattr(L, k)[: int(L.seg_off[-1])])
        l["read_bytes"].append(L.read_bytes[: int(L.read_off[-1])]); l["read_quals"].append(L.read_quals[: int(L.read_off[-1])])
        l["read_off"].append(L.read_off[:-1] + np.uint64(o["lrb"]))
        mo, vo = np.asarray(d["map_off"], np.uint64), np.asarray(d["var_off"], np.uint64)
        map_off.append(mo[1:] + np.uint64(o["nm"])); var_off.append(vo[1:] + np.uint64(o["nvar"]))
        map_group.append(np.asarray(d["map_group"], np.uint32)); n_groups.append(np.asarray(d["n_groups"], np.uint32))
        het_base.append(np.asarray(d["wfa_het_base"], np.uint32) + np.uint32(o["wnv"]))
        vtypes.extend(vt)
        o["ref"] += len(W.reference); o["wab"] += len(W.allele_bytes); o["wnv"] += len(W.position); o["wrb"] += int(W.read_off[-1])
        o["lab"] += len(L.allele_bytes); o["lnv"] += len(L.position); o["lseg"] += int(L.seg_off[-1]); o["lrb"] += int(L.read_off[-1])
        o["nm"] += int(mo[-1]); o["nvar"] += int(vo[-1])
    wtab = {k: cat(v) for k, v in wt.items()}
    ltab = {k: cat(v) for k, v in lt.items()}
    wfa = WfaBatch(wtab, cat(w["reference"]), cat(w["ref_start"]), cat(w["ref_end"]), cat(w["het_lo"]), cat(w["het_hi"]), cat(w["hom_lo"]),
                   cat(w["hom_hi"]), cat(w["read_bytes"]), cat(w["read_off"] + [np.array([o["wrb"]], np.uint64)]))
    local = LocalBatch(ltab, cat(l["var_lo"]), cat(l["var_hi"]), cat(l["read_pos"]), cat(l["seg_off"] + [np.array([o["lseg"]], np.uint64)]),
                       cat(l["seg_ref_start"]), cat(l["seg_read_start"]), cat(l["seg_len"]), cat(l["read_bytes"]), cat(l["read_quals"]),
                       cat(l["read_off"] + [np.array([o["lrb"]], np.uint64)]))
    return dict(wfa=wfa, local=local, map_off=cat(map_off), map_group=cat(map_group), n_groups=cat(n_groups), var_off=cat(var_off),
                wfa_het_base=cat(het_base)), vtypes
