"""astar_solver has no known-answer test in the reference (parity UNPINNED).  These tests pin the
C++ oracle from three independent directions: a separate pure-python restatement, brute-force MEC optima, and
the reference's own invariants."""
import numpy as np
import pytest

import oracle_lib as O
import pyref
from hiphase_b200 import _abi as A
from hiphase_b200 import synth


def _rand_blocks(seed, n, nlo, nhi, p_err=0.05, p_amb=0.05, p_gap=0.03):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        N = int(rng.integers(nlo, nhi + 1))
        R = int(rng.integers(2, 4 * N + 3))
        smax = max(2, min(N, int(rng.integers(2, 14))))
        out.append(synth.gen_block(rng, N, R, lambda r, k, s=smax: r.integers(1, s + 1, k), p_err, p_amb, p_gap, p_ignored=0.08))
    return out


def _check_against_pyref(batch, params, out):
    for b in range(batch.n_blocks):
        h1, h2, st, H = pyref.astar_solver(batch.block(b), params.min_queue_size, params.queue_increment)
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        assert out.status[b] == A.HP_BLOCK_OK
        assert out.h1[v0:v1].tolist() == h1 and out.h2[v0:v1].tolist() == h2, b
        assert {k: int(out.stats[b][k]) for k in st} == st, b
        assert out.heuristic[v0 + b: v1 + b + 1].tolist() == H, b


def test_oracle_vs_python_restatement_default_params():
    batch = A.BlockBatch.from_blocks(_rand_blocks(11, 40, 1, 30))
    out = O.astar_solve(batch)
    assert out.failures == 0
    _check_against_pyref(batch, A.default_params(), out)


def test_oracle_vs_python_restatement_pruning():
    # tiny queue thresholds force min_progress pruning, the threshold snap-back and the full-prune re-keying
    params = A.hp_params(10, 1, 500, 500)
    blocks = _rand_blocks(12, 25, 12, 40, p_err=0.25, p_amb=0.05)
    batch = A.BlockBatch.from_blocks(blocks)
    out = O.astar_solve(batch, params)
    assert out.failures == 0
    assert (out.stats["pruned_solutions"] > 0).any()
    _check_against_pyref(batch, params, out)


def test_oracle_full_prune_path():
    params = A.hp_params(10, 1, 500, 500)     # max_queue = 100: the "full prune" branch (astar_phaser.rs:570-584) fires
    batch = A.BlockBatch.from_blocks(_rand_blocks(13, 15, 30, 50, p_err=0.3))
    out = O.astar_solve(batch, params)
    assert out.failures == 0
    pyref.FULL_PRUNES[0] = 0
    _check_against_pyref(batch, params, out)
    assert pyref.FULL_PRUNES[0] > 0


def test_oracle_matches_bruteforce_mec():
    blocks = _rand_blocks(14, 60, 2, 8, p_err=0.15)
    batch = A.BlockBatch.from_blocks(blocks)
    out = O.astar_solve(batch)
    assert out.failures == 0
    for b, blk in enumerate(blocks):
        assert out.stats[b]["pruned_solutions"] == 0
        assert int(out.stats[b]["actual_cost"]) == synth.brute_force_mec(blk), b


def test_oracle_invariants_and_counters():
    batch = synth.config_c2(n_blocks=6)
    out = O.astar_solve(batch)
    assert out.failures == 0
    for b in range(batch.n_blocks):
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        H = out.heuristic[v0 + b: v1 + b + 1]
        assert H[-1] == 0 and (np.diff(H.astype(np.int64)) <= 0).all()         # astar_phaser.rs:284
        st = out.stats[b]
        assert st["actual_cost"] >= st["estimated_cost"] == H[0]               # phase_stats.rs:163
        assert st["phased_variants"] + st["homozygous_variants"] + st["skipped_variants"] == v1 - v0
        ign = batch.ignored[v0:v1].astype(bool)
        assert (out.h1[v0:v1][ign] == 2).all() and (out.h2[v0:v1][ign] == 2).all()
        assert st["skipped_variants"] == ign.sum()
        c = out.counters[b]
        assert c["evals"] > c["pops"] > 0 and c["cells"] > c["evals"]


def test_oracle_threads_agree():
    batch = synth.config_c2(n_blocks=16)
    a = O.astar_solve(batch, threads=1)
    b = O.astar_solve(batch, threads=4)
    assert (a.h1 == b.h1).all() and (a.h2 == b.h2).all() and (a.stats == b.stats).all()


def test_oracle_rejects_set_ignored_variant():
    blk = {"n_var": 4, "reads": [(0, [0, 1, 0, 1], [5, 5, 5, 5])], "ignored": [0, 1, 0, 0]}
    out = O.astar_solve(A.BlockBatch.from_blocks([blk]))
    assert out.failures == 1 and out.status[0] == A.HP_BLOCK_IGNORED_NOT_NOOVERLAP


def test_singleton_and_empty_reads():
    blocks = [{"n_var": 1, "reads": []}, {"n_var": 3, "reads": []}, {"n_var": 2, "reads": [(0, [0, 1], [9, 9])]}]
    batch = A.BlockBatch.from_blocks(blocks)
    out = O.astar_solve(batch)
    assert out.failures == 0
    _check_against_pyref(batch, A.default_params(), out)
    # no evidence: the first child (0|1) wins every tie -> fully heterozygous
    assert out.h1[:1].tolist() == [0] and out.h2[:1].tolist() == [1]
