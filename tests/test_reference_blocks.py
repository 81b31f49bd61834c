"""A* parity against genuine HiPhase: phase blocks dumped from the Rust binary by tools/reference_dump (inputs of
astar_solver at src/phaser.rs:541 + its AstarResult) are solved by the oracle and by the CUDA path and compared bit for bit.

This image has no Rust toolchain, so tests/golden/reference_blocks/ ships empty and the comparison tests skip; the harness
itself (container reader, `.ref` parser, comparison) is exercised on a fixture written in the same two formats."""
import os

import numpy as np
import pytest

import oracle_lib as O
from helpers import read_reference_result, reference_fixtures, write_reference_result
from hiphase_b200 import lib, synth


def _compare(solve, fixtures):
    assert fixtures
    for hpb, ref in fixtures:
        batch, _ = lib.pack_read_blocks(hpb)
        assert batch.n_blocks == 1
        h1, h2, stats = read_reference_result(ref)
        out = solve(batch)
        assert out.status[0] == 0, hpb
        assert np.array_equal(out.h1, h1) and np.array_equal(out.h2, h2), hpb
        assert np.array_equal(np.array(list(out.stats[0]), np.uint64), stats), (hpb, out.stats[0], stats)


@pytest.fixture
def synthetic_dump(tmp_path):
    """Blocks written in the dump formats (container through hp_pack_write_blocks, results through the `.ref` writer)."""
    batch = synth.config_c3_stream(6, first_block=40)
    ref = O.astar_solve(batch, want_heuristic=False, want_counters=False)
    out = []
    for b in range(batch.n_blocks):
        one = batch.select([b])
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        stem = os.path.join(tmp_path, "block_%08d" % b)
        lib.pack_write_blocks(stem + ".hpb", one, np.arange(v0, v1, dtype=np.int64) * 1000)
        write_reference_result(stem + ".ref", ref.h1[v0:v1], ref.h2[v0:v1], list(ref.stats[b]))
    return reference_fixtures(str(tmp_path))


def test_harness_on_a_synthetic_dump(synthetic_dump):
    assert len(synthetic_dump) == 6
    _compare(lambda b: O.astar_solve(b, want_heuristic=False, want_counters=False), synthetic_dump)
    # a corrupted reference result must be noticed
    hpb, ref = synthetic_dump[0]
    h1, h2, st = read_reference_result(ref)
    write_reference_result(ref, h1, h2, [int(st[0]) + 1] + [int(x) for x in st[1:]])
    with pytest.raises(AssertionError):
        _compare(lambda b: O.astar_solve(b, want_heuristic=False, want_counters=False), synthetic_dump[:1])


def test_oracle_matches_the_rust_binary():
    fx = reference_fixtures()
    if not fx:
        pytest.skip("tests/golden/reference_blocks/ is empty: no Rust toolchain in this image (see tools/reference_dump/README.md)")
    _compare(lambda b: O.astar_solve(b, want_heuristic=False, want_counters=False), fx)


@pytest.mark.gpu
def test_cuda_matches_the_rust_binary(synthetic_dump):
    ctx = lib.Context(device=0)
    _compare(ctx.astar_solve_batch, synthetic_dump)              # the harness against the CUDA path
    fx = reference_fixtures()
    if fx:
        _compare(ctx.astar_solve_batch, fx)
    ctx.close()
    if not fx:
        pytest.skip("tests/golden/reference_blocks/ is empty: no Rust toolchain in this image (see tools/reference_dump/README.md)")


def test_container_layout_written_by_the_rust_patch(tmp_path):
    """dump_block_fixture (tools/reference_dump/hiphase_dump_blocks.patch) restated byte for byte: header, section table,
    64-byte aligned sections in the order of hp_pack_write_blocks.  The C reader must accept it and return the same block."""
    batch = synth.config_c3_stream(1, first_block=7)
    var_pos = np.arange(batch.n_vars, dtype=np.int64) * 37 + 5
    secs = [("var_off", 8, batch.var_off), ("read_off", 8, batch.read_off), ("read_start", 4, batch.read_start),
            ("read_end", 4, batch.read_end), ("cell_off", 8, batch.cell_off), ("alleles", 1, batch.alleles), ("quals", 1, batch.quals),
            ("ignored", 1, batch.ignored), ("is_snv", 1, batch.is_snv), ("var_pos", 8, var_pos)]
    al = lambda x: (x + 63) & ~63
    blob = bytearray(b"HPB200\0\x01") + np.uint32(1).tobytes() + np.uint32(len(secs)).tobytes() + np.uint64(1).tobytes()
    off, offs = al(24 + 40 * len(secs)), []
    for name, es, arr in secs:
        blob += name.encode().ljust(16, b"\0") + np.uint32(es).tobytes() + np.uint32(0).tobytes() + np.uint64(len(arr)).tobytes() + np.uint64(off).tobytes()
        offs.append(off)
        off = al(off + es * len(arr))
    for (name, es, arr), o in zip(secs, offs):
        blob += b"\0" * (o - len(blob))
        blob += np.ascontiguousarray(arr).tobytes()
    path = os.path.join(tmp_path, "block_00000007.hpb")
    open(path, "wb").write(bytes(blob))
    got, vp = lib.pack_read_blocks(path)
    for f in ("var_off", "read_off", "read_start", "read_end", "cell_off", "alleles", "quals", "ignored", "is_snv"):
        assert np.array_equal(getattr(got, f), getattr(batch, f)), f
    assert np.array_equal(vp, var_pos)
