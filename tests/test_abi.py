"""CPU-side checks of the boundary: the shared library loads, exports every symbol of include/hiphase_b200.h,
and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from hiphase_b200 import _abi as A
from hiphase_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    lib.build()
    L = lib.lib()
    header = open(os.path.join(ROOT, "include", "hiphase_b200.h")).read()
    declared = set(re.findall(r"\b(hp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for sym in declared:
        assert getattr(L, sym) is not None
    assert L.hp_abi_version() == 2


def test_struct_sizes_match_header():
    assert C.sizeof(A.hp_phase_stats) == 56 and C.sizeof(A.hp_astar_counters) == 32
    assert C.sizeof(A.hp_params) == 16
    assert C.sizeof(A.hp_block_batch) == 8 + 9 * 8
    p = A.hp_params()
    lib.lib().hp_default_params(C.byref(p))
    assert (p.min_queue_size, p.queue_increment, p.wfa_prune_distance, p.wfa_max_edit_distance) == (1000, 3, 500, 500)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.HiPhaseB200Error) as e:
        lib.Context()
    assert e.value.code == A.HP_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    # the product tree must never import / link / call anything under oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hiphase_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "hp_oracle" not in txt and "oracle_lib" not in txt and "libhp_oracle" not in txt, f


def _c_compile_and_run(tmp_path, source, link=False):
    import subprocess
    src = tmp_path / "probe.c"
    exe = tmp_path / "probe"
    src.write_text(source)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)]
    if link:
        cmd += ["-L", lib.CSRC, "-lhiphase_b200", "-Wl,-rpath," + lib.CSRC]
    subprocess.run(cmd, check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True)


def test_header_is_plain_c_and_the_ctypes_mirrors_match_it(tmp_path):
    """include/hiphase_b200.h compiles as C99 (-pedantic -Werror), and every struct mirrored in _abi.py has the size and the
    field offsets the C compiler gives it (same field names, same order)."""
    structs = [(n, getattr(A, n)) for n in dir(A) if n.startswith("hp_") and isinstance(getattr(A, n), type)
               and issubclass(getattr(A, n), C.Structure)]
    assert len(structs) >= 18
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hiphase_b200.h"', 'int main(void) {']
    for name, cls in structs:
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f[0], name, f[0]))
    lines += ['  return 0;', '}']
    r = _c_compile_and_run(tmp_path, "\n".join(lines))
    assert r.returncode == 0, r.stderr
    got = dict(l.split() for l in r.stdout.strip().splitlines())
    for name, cls in structs:
        assert int(got[name]) == C.sizeof(cls), name
        for f in cls._fields_:
            assert int(got["%s.%s" % (name, f[0])]) == getattr(cls, f[0]).offset, (name, f[0])


def test_c_program_links_the_library_and_gets_the_no_device_error(tmp_path):
    """A C99 caller (what a cc + bindgen build of HiPhase would link): host-only entries work, hp_ctx_create reports
    HP_ERR_NO_DEVICE without a GPU instead of falling back to anything."""
    import torch
    lib.build()
    src = r'''
#include <stdio.h>
#include <string.h>
#include "hiphase_b200.h"
int main(void) {
    hp_params p;
    uint32_t n_var[5] = {2000, 20, 300, 40, 1000};
    uint64_t n_cells[5] = {60000, 600, 9000, 1200, 30000};
    uint64_t cost[5];
    uint32_t shard[5];
    hp_ctx* ctx = NULL;
    int rc;
    hp_default_params(&p);
    if (hp_abi_version() != HP_ABI_VERSION) return 10;
    if (p.min_queue_size != 1000 || p.queue_increment != 3) return 11;
    if (hp_block_costs(5, n_var, n_cells, cost) != HP_OK) return 12;
    if (hp_lpt_partition(cost, 5, 2, shard) != HP_OK) return 13;
    if (shard[0] == shard[4]) return 14;                 /* the two heaviest blocks go to different shards */
    rc = hp_ctx_create(&p, 0, &ctx);
    printf("%d %s\n", rc, hp_last_error(NULL));
    if (rc == HP_OK) hp_ctx_destroy(ctx);
    return 0;
}
'''
    r = _c_compile_and_run(tmp_path, src, link=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    rc = int(r.stdout.split()[0])
    if torch.cuda.is_available():
        assert rc == A.HP_OK
    else:
        assert rc == A.HP_ERR_NO_DEVICE and "no CPU fallback" in r.stdout


def _check_rust_structs(text, at_least):
    found = re.findall(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (\w+)\s*\{(.*?)\}", text, flags=re.S)
    found = [(n, b) for n, b in found if "_private" not in b]
    assert len(found) >= at_least, [n for n, _ in found]
    prim = {"u8": 1, "i8": 1, "u16": 2, "u32": 4, "i32": 4, "c_int": 4, "f32": 4, "u64": 8, "i64": 8, "usize": 8, "f64": 8}

    def size_align(ty):
        ty = ty.strip()
        if ty.startswith("*"):
            return 8, 8
        if ty in prim:
            return prim[ty], prim[ty]
        cls = getattr(A, ty)                       # a nested struct of the header
        return C.sizeof(cls), C.alignment(cls)
    for name, body in found:
        cls = getattr(A, name)
        fields = re.findall(r"pub (\w+)\s*:\s*([^,]+?)\s*(?:,|$)", body.strip(), flags=re.S)
        assert [f for f, _ in fields] == [f[0] for f in cls._fields_], name
        off = 0
        for fname, ty in fields:
            sz, al = size_align(ty)
            off = (off + al - 1) // al * al
            assert off == getattr(cls, fname).offset, (name, fname, ty)
            assert sz == getattr(cls, fname).size, (name, fname, ty)
            off += sz
    return [n for n, _ in found]


def test_rust_structs_of_the_integration_guide_match_the_header():
    """The Rust shim in INTEGRATION.md cannot be compiled here (no cargo), so its #[repr(C)] structs are at least held against the
    header: same field names in the same order, and the offsets repr(C) gives those Rust types equal the C compiler's (through
    the ctypes mirrors, which the test above pins to the header)."""
    _check_rust_structs(open(os.path.join(ROOT, "INTEGRATION.md")).read(), 6)


def test_generated_rust_bindings_are_current_and_complete():
    """bindings/rust/b200_ffi.rs (tools/gen_rust_ffi.py: what bindgen would emit from the header) is up to date, declares every
    exported function and every struct of the header, and its struct layouts equal the C compiler's."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(ROOT, "bindings", "rust", "b200_ffi.rs")).read()
    assert text == gen.generate(open(gen.HEADER).read()), "run python tools/gen_rust_ffi.py"
    assert set(re.findall(r"pub fn (hp_\w+)\(", text)) == set(lib.EXPORTS)
    names = _check_rust_structs(text, 18)
    mirrored = {n for n in dir(A) if n.startswith("hp_") and isinstance(getattr(A, n), type) and issubclass(getattr(A, n), C.Structure)}
    assert set(names) == mirrored
    for n, v in re.findall(r"pub const (HP_\w+): \w+ = (-?\d+);", text):
        if hasattr(A, n):
            assert getattr(A, n) == int(v), n
