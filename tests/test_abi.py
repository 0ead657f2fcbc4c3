"""The C-ABI library loads without a GPU and exports every symbol include/hipstr_b200.h declares;
compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import hipstr_b200
from hipstr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hipstr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hipstr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_ctypes_layout():
    lib = capi.load()
    assert b"sm_100a" in lib.hipstr_version()
    # struct sizes the C side expects (x86-64): 4 int32 + int64 + 15 pointers
    assert C.sizeof(capi.AlignBatch) == 16 + 8 + 15 * 8
    assert C.sizeof(capi.ReadsBatch) == 10 * 8 and C.sizeof(capi.GenotypeOut) == 6 * 8


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hipstr_b200.HipstrError) as e:
        hipstr_b200.Context(0)
    assert e.value.status == 1     # HIPSTR_ERR_NO_DEVICE
    lib = capi.load()
    assert lib.hipstr_align_batch_host(None, None, None, None) == 3   # BAD_ARG, nothing computed


def test_product_does_not_touch_the_oracle():
    """Nothing under hipstr_b200/ may import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "hipstr_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh", "Makefile")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle/" not in src and "liboracle" not in src and "checkers" not in src, os.path.join(d, f)


def test_host_ops_pool_reads():
    import numpy as np
    from hipstr_b200.capi import c_i32p, ptr
    lib = capi.load()
    seqs = [b"ACGT", b"GGGA", b"ACGT", b"ACGT", b"TT", b"GGGA"]
    quals = [b"!5I+", b"ABCD", b"I5!+", b"555,", b"##", b"DCBA"]
    off = np.zeros(len(seqs) + 1, np.int32)
    off[1:] = np.cumsum([len(x) for x in seqs])
    bases, q = b"".join(seqs), b"".join(quals)
    pidx = np.zeros(len(seqs), np.int32)
    first = np.zeros(len(seqs), np.int32)
    poff = np.zeros(len(seqs) + 1, np.int32)
    npools = C.c_int32()
    pb, pq = C.create_string_buffer(len(bases)), C.create_string_buffer(len(bases))
    st = lib.hipstr_pool_reads(len(seqs), ptr(off, c_i32p), bases, q, ptr(pidx, c_i32p), C.byref(npools), ptr(first, c_i32p),
                               ptr(poff, c_i32p), pb, pq)
    assert st == 0 and npools.value == 3
    assert list(pidx) == [0, 1, 0, 0, 2, 1] and list(first[:3]) == [0, 1, 4]
    assert pb.raw[:poff[3]] == b"ACGTGGGATT"
    # upper median (sorted[n/2]) per position: pool 0 has 3 members, pool 1 has 2
    assert pq.raw[:4] == bytes([sorted([a, b, c])[1] for a, b, c in zip(b"!5I+", b"I5!+", b"555,")])
    assert pq.raw[4:8] == bytes([max(a, b) for a, b in zip(b"ABCD", b"DCBA")])
    assert pq.raw[8:10] == b"##"


def test_pool_reads_median_qualities_match_sorting():
    """Upper median per position (sorted[n/2] in signed-char order, base_quality.cpp:11-28) for pools of 1..200 members,
    including bytes above 127 -- the rank-based path (<= 64 members) and the selection path agree with plain sorting."""
    import numpy as np
    from hipstr_b200.capi import c_i32p, ptr
    lib = capi.load()
    rng = np.random.default_rng(9)
    for trial in range(30):
        L = int(rng.integers(1, 130))
        sizes = [int(x) for x in rng.choice([1, 2, 3, 4, 7, 10, 24, 25, 33, 64, 65, 200], int(rng.integers(1, 8)))]
        seqs, quals = [], []
        for p, m in enumerate(sizes):
            seq = bytes(int(v) for v in rng.choice(list(b"ACGT"), L)) + bytes([65 + p])   # distinct per pool
            for _ in range(m):
                seqs.append(seq)
                hi = 256 if trial % 3 == 0 else 127
                quals.append(bytes(int(v) for v in rng.integers(33 if hi == 127 else 0, hi, L + 1)))
        order = rng.permutation(len(seqs))
        seqs, quals = [seqs[i] for i in order], [quals[i] for i in order]
        off = np.zeros(len(seqs) + 1, np.int32)
        off[1:] = np.cumsum([len(x) for x in seqs])
        bases, q = b"".join(seqs), b"".join(quals)
        pidx, first, poff = np.zeros(len(seqs), np.int32), np.zeros(len(seqs), np.int32), np.zeros(len(seqs) + 1, np.int32)
        npools = C.c_int32()
        pb, pq = C.create_string_buffer(len(bases)), C.create_string_buffer(len(bases))
        st = lib.hipstr_pool_reads(len(seqs), ptr(off, c_i32p), bases, q, ptr(pidx, c_i32p), C.byref(npools), ptr(first, c_i32p),
                                   ptr(poff, c_i32p), pb, pq)
        assert st == 0 and npools.value == len(sizes)
        for p in range(npools.value):
            rows = np.array([np.frombuffer(quals[r], np.int8) for r in range(len(seqs)) if pidx[r] == p])
            want = np.sort(rows, axis=0)[rows.shape[0] // 2].astype(np.int8).tobytes()
            assert pq.raw[poff[p]:poff[p + 1]] == want, (trial, p, rows.shape)
