"""Shared helpers for the tests: golden-fixture loading and variant-table construction."""
import json
import os

import numpy as np

from hiphase_b200 import _abi as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_VT = {"snv": A.VT_SNV, "insertion": A.VT_INSERTION, "deletion": A.VT_DELETION, "indel": A.VT_INDEL,
       "sv_insertion": A.VT_SV_INSERTION, "sv_deletion": A.VT_SV_DELETION, "tr": A.VT_TANDEM_REPEAT}


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def variant_table(variants):
    """variants: list of dicts {type,pos,ref_len,a0,a1,i0[,ignored]} with *truncated* allele strings."""
    blob = bytearray()
    t = {k: [] for k in ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len",
                         "index_allele0", "vtype", "ignored")}
    for v in variants:
        a0 = v["a0"].encode() if isinstance(v["a0"], str) else bytes(v["a0"])
        a1 = v["a1"].encode() if isinstance(v["a1"], str) else bytes(v["a1"])
        t["position"].append(v["pos"]); t["ref_len"].append(v["ref_len"])
        t["allele0_off"].append(len(blob)); t["allele0_len"].append(len(a0)); blob += a0
        t["allele1_off"].append(len(blob)); t["allele1_len"].append(len(a1)); blob += a1
        t["index_allele0"].append(v.get("i0", 0)); t["vtype"].append(_VT[v["type"]]); t["ignored"].append(v.get("ignored", 0))
    t["allele_bytes"] = np.frombuffer(bytes(blob), np.uint8) if blob else np.zeros(0, np.uint8)
    return t


def wfa_batch_single(reference, hets, homs, ref_start, ref_end, reads):
    """One variant table (hets then homs), one job per read."""
    ref = np.frombuffer(reference.encode() if isinstance(reference, str) else bytes(reference), np.uint8)
    vt = variant_table(list(hets) + list(homs))
    n = len(reads)
    rb = [np.frombuffer(r.encode() if isinstance(r, str) else bytes(r), np.uint8) for r in reads]
    read_off = np.concatenate([[0], np.cumsum([len(x) for x in rb])]) if n else np.zeros(1)
    return A.WfaBatch(vt, ref, [ref_start] * n, [ref_end] * n, [0] * n, [len(hets)] * n, [len(hets)] * n,
                      [len(hets) + len(homs)] * n, np.concatenate(rb) if n and sum(len(x) for x in rb) else np.zeros(0, np.uint8),
                      read_off)


# ---- fixtures dumped from genuine HiPhase (tools/reference_dump) -------------------------------------------------------
REFERENCE_BLOCKS = os.path.join(GOLDEN, "reference_blocks")


def write_reference_result(path, h1, h2, stats7):
    """The `.ref` format of tools/reference_dump/hiphase_dump_blocks.patch (used by the self-test of the harness)."""
    h1 = np.asarray(h1, np.uint8); h2 = np.asarray(h2, np.uint8)
    with open(path, "wb") as f:
        f.write(b"HPREF1\0\0" + np.uint64(len(h1)).tobytes() + h1.tobytes() + h2.tobytes() + np.asarray(stats7, "<u8").tobytes())


def read_reference_result(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"HPREF1\0\0", "not a HiPhase reference dump: " + path
    n = int(np.frombuffer(raw, "<u8", 1, 8)[0])
    assert len(raw) == 16 + 2 * n + 56, "truncated reference dump: " + path
    h1 = np.frombuffer(raw, np.uint8, n, 16); h2 = np.frombuffer(raw, np.uint8, n, 16 + n)
    return h1, h2, np.frombuffer(raw, "<u8", 7, 16 + 2 * n)


def reference_fixtures(directory=REFERENCE_BLOCKS):
    """[(path.hpb, path.ref)] of the blocks dumped from the Rust binary, sorted."""
    if not os.path.isdir(directory):
        return []
    stems = sorted(f[:-4] for f in os.listdir(directory) if f.endswith(".hpb") and os.path.exists(os.path.join(directory, f[:-4] + ".ref")))
    return [(os.path.join(directory, s + ".hpb"), os.path.join(directory, s + ".ref")) for s in stems]
