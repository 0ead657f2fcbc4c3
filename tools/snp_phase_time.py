"""K7 (SNP phasing log-likelihoods) on a configs[1]-sized batch: kernel time from CUDA events on the launching stream,
the whole C-ABI call with host buffers, and one host core running the oracle restatement on a sample.
usage: python tools/snp_phase_time.py [n_entries] [reps]        (needs a GPU)"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hipstr_b200 import capi
from hipstr_b200.capi import SnpPhasing, SnpPhasingStruct, c_f64p, c_i32p, ptr

n_entries = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000     # 1 000 loci x 100 samples x 30 reads
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(3)
n_sets, chrom_len, read_len = 100, 2_000_000, 150


def flat_batch(n):
    """Vectorised construction (SnpPhasing's list interface is too slow for millions of reads): pairs of 150 bp reads,
    CIGAR 150M or 70M2D80M / 70M2I78M, one SNP per ~700 bp and sample."""
    b = SnpPhasing([], [])
    n_alns = 2 * n
    pos = rng.integers(1000, chrom_len - 2000, n).astype(np.int32)
    aln_pos = np.empty(n_alns, np.int32)
    aln_pos[0::2] = pos
    aln_pos[1::2] = pos + rng.integers(160, 450, n).astype(np.int32)
    kind = rng.choice(3, n_alns, p=[0.8, 0.1, 0.1])
    n_ops = np.where(kind == 0, 1, 3)
    cigar_off = np.zeros(n_alns + 1, np.int32)
    cigar_off[1:] = np.cumsum(n_ops)
    types = np.full(cigar_off[-1], ord("M"), np.uint8)
    lens = np.full(cigar_off[-1], read_len, np.int32)
    idx = cigar_off[:-1]
    d, i = kind == 1, kind == 2
    lens[idx[d]] = 70; types[idx[d] + 1] = ord("D"); lens[idx[d] + 1] = 2; lens[idx[d] + 2] = 80
    lens[idx[i]] = 70; types[idx[i] + 1] = ord("I"); lens[idx[i] + 1] = 2; lens[idx[i] + 2] = 78
    aln_end = aln_pos + np.where(kind == 1, 152, np.where(kind == 2, 148, 150)).astype(np.int32)
    b.entry_aln_off = (2 * np.arange(n + 1)).astype(np.int32)
    b.entry_snp_set = rng.integers(0, n_sets, n).astype(np.int32)
    b.aln_pos, b.aln_end = aln_pos, aln_end.astype(np.int32)
    b.aln_seq_off = (read_len * np.arange(n_alns + 1)).astype(np.int32)
    b.bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n_alns * read_len)].copy()
    b.quals = rng.integers(35, 74, n_alns * read_len).astype(np.uint8)
    b.aln_cigar_off, b.cigar_type, b.cigar_len = cigar_off, types, lens
    per_set = chrom_len // 700
    sets = [np.sort(rng.choice(chrom_len, per_set, replace=False)).astype(np.uint32) for _ in range(n_sets)]
    b.set_off = (per_set * np.arange(n_sets + 1)).astype(np.int32)
    b.snp_pos = np.concatenate(sets)
    b.snp_base1 = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(b.snp_pos))].copy()
    b.snp_base2 = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(b.snp_pos))].copy()
    b.n_entries, b.n_alns, b.n_sets = n, n_alns, n_sets
    b.struct = SnpPhasingStruct(n, ptr(b.entry_aln_off, c_i32p), ptr(b.entry_snp_set, c_i32p), n_alns, ptr(b.aln_pos, c_i32p),
                                ptr(b.aln_end, c_i32p), ptr(b.aln_seq_off, c_i32p), b.bases.ctypes.data, b.quals.ctypes.data,
                                ptr(b.aln_cigar_off, c_i32p), b.cigar_type.ctypes.data, ptr(b.cigar_len, c_i32p), n_sets,
                                ptr(b.set_off, c_i32p), b.snp_pos.ctypes.data, b.snp_base1.ctypes.data, b.snp_base2.ctypes.data)
    return b


batch = flat_batch(n_entries)
ctx = capi.Context(0)
ctx.enable_timing(True)
for _ in range(3):
    p1, p2, counts = ctx.snp_phasing(batch)
kernel_ms, call_s = [], []
for _ in range(reps):
    t = time.perf_counter()
    p1, p2, counts = ctx.snp_phasing(batch)
    call_s.append(time.perf_counter() - t)
    kernel_ms.append(float(ctx.lib.hipstr_last_kernel_ms(ctx.h)))
h2d, d2h, launches = ctx.traffic()
# algorithmic bytes per entry: offsets, set, positions, CIGAR of both alignments, the bases + qualities actually under SNPs
# (one 32-byte sector each), a binary search over the set (log2 of its size probes), and the 32 bytes of results
n_snps_hit = int(counts[:, :3].sum())
small = flat_batch(200_000)
f = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")).oracle_snp_phasing
f.restype = C.c_int32
f.argtypes = [C.POINTER(SnpPhasingStruct), c_f64p, c_f64p, c_i32p]
t = time.perf_counter()
st, o1, o2, oc = small.run(f)
cpu_s = time.perf_counter() - t
s1, s2, sc = ctx.snp_phasing(small)
assert st == 0 and np.array_equal(o1.view(np.uint64), s1.view(np.uint64)) and np.array_equal(oc, sc)
k = float(np.median(kernel_ms)) / 1e3
print(json.dumps({"entries": n_entries, "alignments": 2 * n_entries, "snps_overlapped": n_snps_hit, "phased_entries": int((p1 != p2).sum()),
                  "kernel_ms": round(k * 1e3, 3), "entries_per_s_kernel": n_entries / k, "call_ms": round(float(np.median(call_s)) * 1e3, 1),
                  "entries_per_s_call": n_entries / float(np.median(call_s)), "h2d_bytes": h2d, "d2h_bytes": d2h, "launches": launches,
                  "input_GBps_kernel": h2d / k / 1e9, "oracle_one_core_entries_per_s": 200_000 / cpu_s,
                  "sample_checked_against_oracle": True}))
