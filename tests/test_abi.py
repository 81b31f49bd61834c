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
