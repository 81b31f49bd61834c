"""Packed phase-block container and batch runner: hp_pack_* / hp_write_phase_stats / hp_phase_blocks."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from hiphase_b200 import _abi as A
from hiphase_b200 import lib, synth


def _positions(batch, seed=5):
    rng = np.random.default_rng(seed)
    pos = np.zeros(batch.n_vars, np.int64)
    for b in range(batch.n_blocks):
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        pos[v0:v1] = 10_000 * (b + 1) + np.cumsum(rng.integers(1, 2000, v1 - v0))
    return pos


def test_pack_round_trip(tmp_path):
    batch = synth.config_c2(n_blocks=7, n_var=60, n_reads=25)
    pos = _positions(batch)
    for vp in (pos, None):
        path = str(tmp_path / "blocks.hpb")
        lib.pack_write_blocks(path, batch, vp)
        back, vp2 = lib.pack_read_blocks(path)
        for f in A.BlockBatch.FIELDS:
            assert np.array_equal(getattr(back, f), getattr(batch, f)), f
        assert (vp2 is None) == (vp is None) and (vp is None or np.array_equal(vp2, vp))
    # every section is 64-byte aligned and the file starts with the magic
    raw = open(path, "rb").read()
    assert raw[:8] == b"HPB200\x00\x01"
    n_sections = int.from_bytes(raw[12:16], "little")
    for i in range(n_sections):
        off = int.from_bytes(raw[24 + 40 * i + 32: 24 + 40 * i + 40], "little")
        assert off % 64 == 0


def test_pack_empty_batch(tmp_path):
    empty = A.BlockBatch([0], [0], [], [], [0], [], [], [], [])
    path = str(tmp_path / "empty.hpb")
    lib.pack_write_blocks(path, empty)
    back, vp = lib.pack_read_blocks(path)
    assert back.n_blocks == 0 and vp is None


def test_pack_rejects_corrupt_files(tmp_path):
    batch = synth.config_c2(n_blocks=3, n_var=30, n_reads=12)
    path = str(tmp_path / "blocks.hpb")
    lib.pack_write_blocks(path, batch, _positions(batch))
    raw = bytearray(open(path, "rb").read())
    cases = {"magic": bytes([0]) + bytes(raw[1:]), "truncated": bytes(raw[: len(raw) // 2]), "short": bytes(raw[:10])}
    bad_count = bytearray(raw); bad_count[24 + 24] ^= 1          # count of the first section
    cases["count"] = bytes(bad_count)
    for name, data in cases.items():
        p = str(tmp_path / (name + ".hpb"))
        open(p, "wb").write(data)
        with pytest.raises(lib.HiPhaseB200Error):
            lib.pack_read_blocks(p)
    with pytest.raises(lib.HiPhaseB200Error):
        lib.pack_read_blocks(str(tmp_path / "missing.hpb"))


def test_stats_writer_columns(tmp_path):
    batch = synth.config_c2(n_blocks=4, n_var=40, n_reads=16)
    pos = _positions(batch)
    ref = O.astar_solve(batch)                                   # any filled AstarOut will do for the writer
    for ext, delim in ((".tsv", "\t"), (".csv", ",")):
        path = str(tmp_path / ("stats" + ext))
        lib.write_phase_stats(path, batch, ref, pos, first_block_index=10)
        rows = [l.rstrip("\n").split(delim) for l in open(path)]
        assert rows[0] == ["block_index", "start", "end", "num_variants", "num_reads", "pruned_solutions", "estimated_cost", "actual_cost",
                           "cost_ratio", "phased_variants", "homozygous_variants", "skipped_variants", "solver_status"]
        assert len(rows) == 1 + batch.n_blocks
        for b, r in enumerate(rows[1:]):
            v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
            st = ref.stats[b]
            assert [int(x) for x in r[:8]] == [10 + b, pos[v0], pos[v1 - 1], v1 - v0, int(batch.read_off[b + 1] - batch.read_off[b]),
                                               int(st["pruned_solutions"]), int(st["estimated_cost"]), int(st["actual_cost"])]
            exp = 1.0 if st["actual_cost"] == 0 else float(st["estimated_cost"]) / float(st["actual_cost"])      # phase_stats.rs:184-198
            assert float(r[8]) == exp
            assert [int(x) for x in r[9:12]] == [int(st["phased_variants"]), int(st["homozygous_variants"]), int(st["skipped_variants"])]


@pytest.mark.gpu
def test_runner_matches_api(tmp_path):
    """hp_phase_blocks on a packed file == the C ABI called directly == the oracle."""
    batch = synth.config_c2(n_blocks=24, n_var=80, n_reads=30)
    pos = _positions(batch)
    path = str(tmp_path / "blocks.hpb")
    lib.pack_write_blocks(path, batch, pos)
    assert os.path.exists(lib.RUNNER_PATH), "hp_phase_blocks not built (run __graft_entry__.build())"
    prefix = str(tmp_path / "run")
    r = subprocess.run([lib.RUNNER_PATH, path, prefix, "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = O.astar_solve(batch)
    post = O.post_solve(batch, pos, ref.h1, ref.h2)
    rows = [l.rstrip("\n").split("\t") for l in open(prefix + ".stats.tsv")][1:]
    for b, row in enumerate(rows):
        st = ref.stats[b]
        assert [int(row[5]), int(row[6]), int(row[7]), int(row[9]), int(row[10]), int(row[11]), int(row[12])] == \
               [int(st["pruned_solutions"]), int(st["estimated_cost"]), int(st["actual_cost"]), int(st["phased_variants"]),
                int(st["homozygous_variants"]), int(st["skipped_variants"]), 0]
    haps = np.loadtxt(prefix + ".haps.tsv", dtype=np.int64, skiprows=1)
    assert haps.shape[0] == batch.n_vars
    assert np.array_equal(haps[:, 2], pos) and np.array_equal(haps[:, 3], ref.h1) and np.array_equal(haps[:, 4], ref.h2)
    assert np.array_equal(haps[:, 5].astype(np.uint64), post.block_tags)


def test_runner_fails_loudly_without_gpu(tmp_path):
    """The batch runner has no CPU solver inside: without a CUDA device it reports the library's error and exits non-zero."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib.build()
    batch = synth.config_c2(n_blocks=2, n_var=30, n_reads=12)
    path = str(tmp_path / "blocks.hpb")
    lib.pack_write_blocks(path, batch, _positions(batch))
    r = subprocess.run([lib.RUNNER_PATH, path, str(tmp_path / "run"), "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "error" in r.stderr.lower()
    assert not os.path.exists(str(tmp_path / "run.stats.tsv"))


def test_corrupt_counts_are_rejected_not_followed(tmp_path):
    """A section count that makes elem_size * count wrap, or a read region outside its block, is an invalid file."""
    import struct
    from hiphase_b200 import synth
    batch = synth.config_c3_stream(2, first_block=3)
    path = os.path.join(tmp_path, "ok.hpb")
    lib.pack_write_blocks(path, batch)
    raw = bytearray(open(path, "rb").read())
    n_sections = struct.unpack_from("<I", raw, 12)[0]
    for i in range(n_sections):                     # read_start: count 2^62 (4 * count wraps to 0)
        name = bytes(raw[24 + 40 * i: 24 + 40 * i + 16]).rstrip(b"\0")
        if name == b"read_start":
            bad = bytearray(raw)
            struct.pack_into("<Q", bad, 24 + 40 * i + 24, 1 << 62)
            p2 = os.path.join(tmp_path, "wrap.hpb"); open(p2, "wb").write(bytes(bad))
            with pytest.raises(lib.HiPhaseB200Error):
                lib.pack_read_blocks(p2)
        if name == b"read_end":                     # first read ends beyond its block's variants
            off = struct.unpack_from("<Q", raw, 24 + 40 * i + 32)[0]
            bad = bytearray(raw)
            struct.pack_into("<I", bad, off, 1 << 30)
            p3 = os.path.join(tmp_path, "region.hpb"); open(p3, "wb").write(bytes(bad))
            with pytest.raises(lib.HiPhaseB200Error):
                lib.pack_read_blocks(p3)
