"""Slow, independent pure-Python restatement of the A* phaser (src/astar_phaser.rs:246-633).

Second opinion for the C++ oracle on the three functions the reference leaves unpinned (astar_subsolver,
calculate_astar_heuristic, astar_solver).  Written separately from oracle/astar_oracle.cpp on purpose:
heapq with tuple keys, dict-free linear read scans, haplotypes as tuples.  Small inputs only.
"""
import heapq

ORDER = ((0, 1), (1, 0), (0, 0), (1, 1))
FULL_PRUNES = [0]     # how many times the full-prune re-keying fired (astar_phaser.rs:570-584)


def _score(read, hap, offset):
    start, al, qu = read
    end = start + len(al)
    if len(hap) + offset <= start or offset >= end:
        return 0
    s = 0
    for i in range(max(start, offset), min(end, offset + len(hap))):
        h = hap[i - offset]
        if h < 2 and al[i - start] != h:
            s += int(qu[i - start])
    return s


def _extend(parent, a1, a2, heur, reads, offset, idx):
    frozen, _fluid, _h, h1, h2, hets, _ = parent
    h1, h2 = h1 + (a1,), h2 + (a2,)
    hap_len = len(h1) + offset
    fluid = 0
    for rd in reads:
        start, al, _ = rd
        end = start + len(al)
        if start < hap_len and end > hap_len - 1:
            c = min(_score(rd, h1, offset), _score(rd, h2, offset))
            if end <= hap_len:
                frozen += c
            else:
                fluid += c
    return (frozen, fluid, heur, h1, h2, hets + (a1 != a2), idx)


def _key(n):
    return (n[0] + n[1] + n[2], -n[5], n[6])


def subsolver(offset, size, reads, H, bad, min_q, inc):
    assert H[offset] == 0
    root = (0, 0, H[offset + 1], (), (), 0, 0)
    pq = [(_key(root), root)]
    nxt, next_expected, max_cost, visits = 1, 0, 0, 0
    max_visits = min_q + inc * size
    while len(pq[0][1][3]) < size and visits < max_visits:
        _, top = heapq.heappop(pq)
        L = len(top[3])
        visits += 1
        if L == next_expected:
            max_cost = max(max_cost, top[0] + top[1] + top[2])
            next_expected += 1
        if bad[offset + L]:
            kids = [(2, 2)]
        else:
            kids = [o for o in ORDER if not (o == (1, 0) and top[3] == top[4])]
        for a1, a2 in kids:
            n = _extend(top, a1, a2, H[offset + L + 1], reads, offset, nxt)
            nxt += 1
            heapq.heappush(pq, (_key(n), n))
    if len(pq[0][1][3]) == size:
        t = pq[0][1]
        max_cost = max(max_cost, t[0] + t[1] + t[2])
        next_expected += 1
    return max_cost, next_expected - 1


def heuristic(N, reads, bad, min_q, inc, max_seg=40):
    H = [0] * (N + 1)
    clip = 1
    for v in range(N - 1, -1, -1):
        est, solved = subsolver(v, clip, reads, H, bad, min_q // 10, inc)
        assert solved >= min(clip, 2)
        if bad[v]:
            H[v] = H[v + 1]
        else:
            assert est >= H[v + 1]
            H[v] = est
        clip = min(solved + 1, max_seg)
    return H


def astar_solver(block, min_q=1000, inc=3):
    N = block["n_var"]
    reads = [(int(s), [int(x) for x in a], [int(x) for x in q]) for (s, a, q) in block["reads"]]
    bad = [bool(x) for x in block["ignored"]]
    is_snv = block["is_snv"]
    H = heuristic(N, reads, bad, min_q, inc)
    thresh, max_queue, min_progress, next_expected, pruned = min_q, 10 * min_q, 0, 0, 0
    counts = [0] * (N + 1)
    tracked = [0]        # entries with len >= min_progress

    root = (0, 0, H[0], (), (), 0, 0)
    pq = [(_key(root), root)]
    counts[0] += 1
    tracked[0] += 1
    nxt = 1
    while len(pq[0][1][3]) < N:
        _, top = heapq.heappop(pq)
        L = len(top[3])
        counts[L] -= 1
        if L >= min_progress:
            tracked[0] -= 1
        if L == next_expected:
            next_expected += 1
            if pruned == 0:
                thresh += inc
        if L < min_progress:
            if pruned == 0:
                thresh = min_q
            pruned += 1
            continue
        kids = [(2, 2)] if bad[L] else [o for o in ORDER if not (o == (1, 0) and top[3] == top[4])]
        for a1, a2 in kids:
            n = _extend(top, a1, a2, H[L + 1], reads, 0, nxt)
            nxt += 1
            heapq.heappush(pq, (_key(n), n))
            counts[L + 1] += 1
            tracked[0] += 1          # L + 1 > min_progress always holds here
        while tracked[0] > thresh and min_progress < next_expected:
            tracked[0] -= counts[min_progress]
            min_progress += 1
            if len(pq) > max_queue:
                FULL_PRUNES[0] += 1
                pq = [((0, k[1], k[2]), n) if len(n[3]) < min_progress else (k, n) for (k, n) in pq]
                heapq.heapify(pq)
    _, top = heapq.heappop(pq)
    h1, h2 = list(top[3]), list(top[4])
    st = {"pruned_solutions": pruned, "estimated_cost": H[0], "actual_cost": top[0] + top[1] + top[2],
          "phased_variants": 0, "phased_snvs": 0, "homozygous_variants": 0, "skipped_variants": 0}
    for i in range(N):
        if h1[i] != h2[i]:
            st["phased_variants"] += 1
            st["phased_snvs"] += int(bool(is_snv[i]))
        elif h1[i] == 2:
            st["skipped_variants"] += 1
        else:
            st["homozygous_variants"] += 1
    return h1, h2, st, H


# ---- independent restatement of local_realignment (src/read_parsing.rs:121-503) in plain Python ----------------------
def py_edit_distance(a, b):
    """sequence_alignment.rs:6-38"""
    prev = list(range(len(a) + 1))
    for i, cb in enumerate(b):
        row = [i + 1] + [0] * len(a)
        for j, ca in enumerate(a):
            row[j + 1] = min(prev[j + 1] + 1, row[j] + 1, prev[j] + (0 if ca == cb else 1))
        prev = row
    return prev[len(a)]


def _as_u8(f):
    if f != f or f <= 0.0:
        return 0
    return 255 if f >= 255.0 else int(f)


def _rmin(a, b):      # f64::min: the non-NaN operand wins
    return b if a != a else (a if b != b else min(a, b))


def _rmax(a, b):
    return b if a != a else (a if b != b else max(a, b))


def py_local_realignment(variants, read_pos, segs, seq, quals):
    """variants: list of dict(pos, ref_len, prefix_len, postfix_len, a0, a1 (bytes, full alleles), vtype, ignored);
    segs: [(ref_start, read_start, len)].  Returns (alleles, quals, match_class, status)."""
    lookup, max_position = {}, read_pos
    for (r, d, n) in segs:
        for i in range(n):
            lookup[r + i] = d + i
            max_position = max(max_position, r + i)
    lo, hi = read_pos, max_position + 1
    out_a, out_q, out_c, status = [], [], [], 0
    last_deletion_end = 0
    for v in variants:
        pos, vt = v["pos"], v["vtype"]
        allele, qual, exact, overlaps = 3, 0, False, False
        if v["ignored"]:
            pass
        elif pos < last_deletion_end:
            allele, overlaps = 2, True
        elif vt in (0, 1, 2, 3, 4, 9):
            pl, ql, rl = v["prefix_len"], v["postfix_len"], v["ref_len"]
            first_start, last_start, first_end, last_end = pos - pl, pos + 1, pos + rl, pos + rl + ql + 1
            cs = next((lookup[c] for c in range(last_start - 1, first_start - 1, -1) if c in lookup), None)
            ce = next((lookup[c] for c in range(first_end, last_end) if c in lookup), None)
            ss = se = None
            start_clip = end_clip = 0
            if cs is not None and ce is not None:
                for sc in range(first_start, last_start):
                    start_clip += 1
                    if sc in lookup:
                        if cs - lookup[sc] > 2 * pl:
                            continue
                        ss = lookup[sc]
                        for ec in range(last_end - 1, first_end - 1, -1):
                            end_clip += 1
                            if ec in lookup:
                                if lookup[ec] - ce > 2 * ql:
                                    continue
                                se = lookup[ec]
                                break
                        break
            if ss is not None:
                overlaps = True
                if se is not None:
                    obs = bytes(seq[ss:se])
                    if obs == v["a0"]:
                        allele, exact = 0, True
                    elif obs == v["a1"]:
                        allele, exact = 1, True
                    else:
                        h, t = start_clip - 1, end_clip - 1
                        d0 = py_edit_distance(obs, v["a0"][h:len(v["a0"]) - t])
                        d1 = py_edit_distance(obs, v["a1"][h:len(v["a1"]) - t])
                        allele = 0 if d0 < d1 else (1 if d0 > d1 else 2)
                    s = 0.0
                    for q in quals[ss:se]:
                        s += (1.0 / float(q)) if q else float("inf")
                    n = float(se - ss)
                    harmonic = (n / s) if s != 0.0 else float("nan")     # 0/0 (empty slice)
                    factor = _rmin(harmonic / 40.0, 1.0)
                    base = {0: 80.0, 9: 40.0, 4: 20.0}.get(vt, 10.0)
                    qual = _as_u8(_rmax(base * factor, 1.0))
                else:
                    allele = 2
            elif lo <= pos < hi:
                allele, overlaps = 2, True
        elif vt == 5:
            if lo <= pos < hi:
                overlaps, allele = True, 2
                last_start, first_end = pos + 1, pos + v["ref_len"]
                if lo <= first_end < hi:
                    expected = first_end - last_start
                    sa = last_start
                    while sa not in lookup:
                        if sa <= lo:
                            break
                        sa -= 1
                    ea = first_end
                    while ea not in lookup:
                        ea += 1
                        if ea >= hi:
                            break
                    deleted = sum(1 for c in range(sa, ea) if c not in lookup)
                    ratio = (deleted / expected) if expected else (float("nan") if deleted == 0 else float("inf"))
                    if ratio < 0.33:
                        allele, qual, exact = 0, _as_u8(_rmax(20.0 * (1.0 - ratio), 1.0)), ratio == 0.0
                    elif abs(1.0 - ratio) < 0.33:
                        allele, qual, exact = 1, _as_u8(_rmax(20.0 * (1.0 - abs(1.0 - ratio)), 1.0)), ratio == 1.0
                        last_deletion_end = first_end
        else:
            status = 1
        out_a.append(allele); out_q.append(qual); out_c.append((1 if overlaps else 0) | (2 if exact else 0))
    return out_a, out_q, out_c, status
