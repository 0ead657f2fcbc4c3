"""K7 -- SNP phasing log-likelihoods (hipstr_snp_phasing_batch_host) against the oracle and the UNMODIFIED reference
(calc_het_snp_factors over a real SNPTree, src/snp_phasing_quality.cpp + src/snp_tree.h, via oracle/ref_bam_harness.cpp).
Doubles are compared bit for bit: the sums are additions of host-computed table entries in the reference's order."""
import ctypes as C

import numpy as np
import pytest

import checkers
from hipstr_b200 import capi
from hipstr_b200.capi import SnpPhasing, SnpPhasingStruct, c_f64p, c_i32p, ptr

needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")


def random_alignment(rng, lo=100, hi=1200):
    pos = int(rng.integers(lo, hi))
    ops = []
    if rng.random() < 0.1:
        ops.append(("H", int(rng.integers(1, 20))))
    if rng.random() < 0.3:
        ops.append(("S", int(rng.integers(1, 25))))
    n_core = int(rng.integers(1, 7))
    for i in range(n_core):
        kinds = "M" * 6 + "=X" + ("ID" * 2 if 0 < i < n_core - 1 else "")
        t = kinds[int(rng.integers(0, len(kinds)))]
        if ops and ops[-1][0] == t:
            t = "M" if t != "M" else "="
        ops.append((t, int(rng.integers(1, 60 if t in "M=X" else 8))))
    if rng.random() < 0.3:
        ops.append(("S", int(rng.integers(1, 25))))
    if rng.random() < 0.1:
        ops.append(("H", int(rng.integers(1, 20))))
    n_bases = sum(n for t, n in ops if t in "SM=XI")
    end = pos + sum(n for t, n in ops if t in "M=XD")
    bases = "".join("ACGTN"[int(x)] for x in rng.choice(5, n_bases, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
    span = (33, 75) if rng.random() < 0.8 else (1, 256)      # also bytes below '!' / above 'J' / above 127
    quals = bytes(int(x) for x in rng.integers(span[0], span[1], n_bases))
    return (pos, end, bases, quals, ops)


def random_batch(rng, n_entries, n_sets=4, snps_per_set=120, min_snps=0):
    sets = []
    for _ in range(n_sets):
        positions = np.sort(rng.choice(np.arange(50, 1500), int(rng.integers(min_snps, snps_per_set)), replace=False))
        sets.append([(int(p), "ACGT"[int(rng.integers(0, 4))], "ACGT"[int(rng.integers(0, 4))]) for p in positions])
    entries = []
    for _ in range(n_entries):
        st = int(rng.integers(-1, n_sets))
        entries.append((st, [random_alignment(rng) for _ in range(1 + int(rng.random() < 0.5))]))
    return SnpPhasing(entries, sets)


def run_oracle(batch):
    f = checkers.oracle().oracle_snp_phasing
    f.restype = C.c_int32
    f.argtypes = [C.POINTER(SnpPhasingStruct), c_f64p, c_f64p, c_i32p]
    st, p1, p2, counts = batch.run(f)
    assert st == 0
    return p1, p2, counts


def run_ref(b):
    f = checkers.ref().ref_snp_phasing
    f.restype = None
    p1, p2 = np.zeros(b.n_entries), np.zeros(b.n_entries)
    counts = np.zeros((b.n_entries, 2), np.int32)
    vp = C.c_void_p
    f.argtypes = [C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, vp, vp, c_i32p, vp, c_i32p, C.c_int32, c_i32p, vp, vp, vp,
                  c_f64p, c_f64p, c_i32p]
    f(b.n_entries, ptr(b.entry_aln_off, c_i32p), ptr(b.entry_snp_set, c_i32p), ptr(b.aln_pos, c_i32p), ptr(b.aln_end, c_i32p),
      ptr(b.aln_seq_off, c_i32p), b.bases.ctypes.data, b.quals.ctypes.data, ptr(b.aln_cigar_off, c_i32p), b.cigar_type.ctypes.data,
      ptr(b.cigar_len, c_i32p), b.n_sets, ptr(b.set_off, c_i32p), b.snp_pos.ctypes.data, b.snp_base1.ctypes.data,
      b.snp_base2.ctypes.data, ptr(p1, c_f64p), ptr(p2, c_f64p), ptr(counts, c_i32p))
    return p1, p2, counts


def same_bits(a, b):
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_reference(seed):
    batch = random_batch(np.random.default_rng(seed), 600)
    p1, p2, counts = run_oracle(batch)
    r1, r2, rc = run_ref(batch)
    assert same_bits(p1, r1) and same_bits(p2, r2)
    assert np.array_equal(counts[:, 0] + counts[:, 1], rc[:, 0]) and np.array_equal(counts[:, 2], rc[:, 1])
    assert (p1 != p2).sum() > 50 and (counts[:, 2] > 0).sum() > 20     # the cases exercise all three branches


@needs_ref
def test_reference_tree_with_many_snps_is_a_range_query():
    """SNPTree only splits above 64 SNPs (snp_tree.h:72): a large set checks that findContained is the sorted range query
    the flat SNP sets assume."""
    rng = np.random.default_rng(11)
    batch = random_batch(rng, 400, n_sets=2, snps_per_set=1400, min_snps=700)
    assert max(np.diff(batch.set_off)) > 600
    p1, p2, counts = run_oracle(batch)
    r1, r2, rc = run_ref(batch)
    assert same_bits(p1, r1) and same_bits(p2, r2)
    assert np.array_equal(counts[:, 0] + counts[:, 1], rc[:, 0]) and np.array_equal(counts[:, 2], rc[:, 1])


def duplicate_pos_batch():
    """Two VCF rows at one POS (a multi-allelic site split by `bcftools norm -m-`): 1|0 on one row, 0|1 on the other.
    create_snp_trees emits both SNPs; the set stays small so that the reference's std::sort keeps their order."""
    read = (100, 120, "ACGTACGTACGTACGTACGT", b"5" * 20, [("M", 20)])
    mate = (300, 310, "TTTTTTTTTT", b"I" * 10, [("M", 10)])
    sets = [[(103, "T", "G"), (108, "A", "C"), (108, "C", "A"), (108, "G", "T"), (115, "T", "A"), (304, "T", "A"), (304, "A", "T")]]
    return SnpPhasing([(0, [read, mate]), (0, [read])], sets)


@needs_ref
def test_duplicate_snp_positions_match_reference():
    batch = duplicate_pos_batch()
    p1, p2, counts = run_oracle(batch)
    r1, r2, rc = run_ref(batch)
    assert same_bits(p1, r1) and same_bits(p2, r2)
    assert np.array_equal(counts[:, 0] + counts[:, 1], rc[:, 0]) and np.array_equal(counts[:, 2], rc[:, 1])
    assert counts[0].tolist()[:3] == [4, 2, 1]


def test_oracle_hand_case():
    """One read 10M2D5M1I4M at 100 with SNPs under a match, inside the deletion, after the insertion and in the mate."""
    read = (100, 121, "ACGTACGTACGTACGTACGT", b"5" * 20, [("M", 10), ("D", 2), ("M", 5), ("I", 1), ("M", 4)])
    mate = (300, 310, "TTTTTTTTTT", b"I" * 10, [("S", 2), ("M", 8)])
    sets = [[(99, "A", "C"), (103, "T", "G"), (110, "A", "C"), (117, "A", "C"), (120, "T", "G"), (121, "A", "C"), (302, "T", "A")]]
    p1, p2, counts = run_oracle(SnpPhasing([(0, [read, mate])], sets))
    ok = lambda q: np.log(1.0 - 10.0 ** ((q - 33) / -10.0))
    bad = lambda q: np.log(10.0 ** ((q - 33) / -10.0) / 3.0)
    # 103 -> read base 3 'T' (haplotype one); 110 in the deletion; 117 -> base 15+1 (after the insertion) = 'A' -> one;
    # 120 -> base 19 'T' -> one; 302 -> mate base 2+2 'T' -> one
    want1 = ((ok(53) + ok(53)) + ok(53)) + ok(73)
    want2 = ((bad(53) + bad(53)) + bad(53)) + bad(73)
    assert p1[0] == want1 and p2[0] == want2
    assert counts[0].tolist() == [4, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n", [(5, 1), (6, 700), (7, 20000)])
def test_kernel_matches_oracle(seed, n):
    batch = random_batch(np.random.default_rng(seed), n)
    with capi.Context() as ctx:
        p1, p2, counts = ctx.snp_phasing(batch)
        assert ctx.traffic()[2] == 1
    o1, o2, oc = run_oracle(batch)
    assert same_bits(p1, o1) and same_bits(p2, o2) and np.array_equal(counts, oc)


@pytest.mark.gpu
def test_kernel_duplicate_snp_positions():
    batch = duplicate_pos_batch()
    with capi.Context() as ctx:
        p1, p2, counts = ctx.snp_phasing(batch)
    o1, o2, oc = run_oracle(batch)
    assert same_bits(p1, o1) and same_bits(p2, o2) and np.array_equal(counts, oc)


@pytest.mark.gpu
@needs_ref
def test_kernel_matches_reference():
    batch = random_batch(np.random.default_rng(21), 1500, n_sets=3, snps_per_set=900, min_snps=300)
    with capi.Context() as ctx:
        p1, p2, counts = ctx.snp_phasing(batch)
    r1, r2, rc = run_ref(batch)
    assert same_bits(p1, r1) and same_bits(p2, r2)
    assert np.array_equal(counts[:, 0] + counts[:, 1], rc[:, 0]) and np.array_equal(counts[:, 2], rc[:, 1])


@pytest.mark.gpu
def test_kernel_edge_cases():
    with capi.Context() as ctx:
        # nothing to do
        p1, p2, counts = ctx.snp_phasing(SnpPhasing([], []))
        assert len(p1) == 0
        # samples without SNP information, and an empty SNP set
        aln = (100, 110, "ACGTACGTAC", b"I" * 10, [("M", 10)])
        p1, p2, counts = ctx.snp_phasing(SnpPhasing([(-1, [aln]), (0, [aln])], [[]]))
        assert p1.tolist() == [0, 0] and p2.tolist() == [0, 0] and not counts.any()
        # a CIGAR character the reference dies on, and SNP positions out of order
        bad = (100, 110, "ACGTACGTAC", b"I" * 10, [("P", 2), ("M", 10)])
        with pytest.raises(RuntimeError):
            ctx.snp_phasing(SnpPhasing([(0, [bad])], [[(104, "A", "C")]]))
        with pytest.raises(RuntimeError):
            ctx.snp_phasing(SnpPhasing([(0, [aln])], [[(104, "A", "C"), (103, "A", "C")]]))
        # CIGAR shorter than the span the alignment claims
        short = (100, 130, "ACGTACGTAC", b"I" * 10, [("M", 10)])
        with pytest.raises(RuntimeError):
            ctx.snp_phasing(SnpPhasing([(0, [short])], [[(120, "A", "C")]]))
