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
