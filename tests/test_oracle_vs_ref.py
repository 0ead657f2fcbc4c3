"""The CPU restatement (oracle/hipstr_oracle.cpp) against the UNMODIFIED reference sources compiled
into oracle/_ref/libhipstr_ref.so (see oracle/Makefile).  This is what pins the oracle: the reference
ships no golden vectors for the path.  Runs wherever the reference library exists
(the build container, and the GPU box, which receives the prebuilt .so); the golden fixtures in
tests/golden/ (test_golden.py) cover the same ground where it does not."""
import ctypes as C

import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import BatchBuilder, c_i32p, ptr

ref = checkers.ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref/libhipstr_ref.so not built")


@needs_ref
def test_fast_lse_bit_exact():
    o = checkers.oracle()
    rng = np.random.default_rng(0)
    a = rng.uniform(-50, 0, 20000)
    b = a + rng.uniform(-9, 9, 20000)
    for x, y in zip(a, b):
        assert o.oracle_fast_lse2(x, y) == ref.ref_fast_lse2(x, y)
    for _ in range(3000):
        n = int(rng.integers(1, 40))
        v = np.ascontiguousarray(rng.uniform(-30, 0, n) * rng.choice([1.0, 0.1, 10.0]))
        p = v.ctypes.data_as(C.POINTER(C.c_double))
        assert o.oracle_fast_lse_vec(p, n) == ref.ref_fast_lse_vec(p, n)


@needs_ref
def test_gray_code_order_matches_reference_iterator():
    o = checkers.oracle()
    # the reference's Haplotype constructor asserts exactly 3 blocks (Haplotype.cpp:9)
    for nopts in ([2, 3, 2], [1, 5, 1], [3, 1, 4], [2, 2, 2], [4, 3, 3]):
        bb = BatchBuilder()
        blocks = []
        for i, k in enumerate(nopts):
            per = 2 if (i % 2 == 1) else 0
            blocks.append((per, ["ACGTAC" + "AC" * j for j in range(k)]))
        blocks[-1] = (0, blocks[-1][1])
        bb.add_locus(blocks, [("ACGTACGTAC", "IIIIIIIIII", 4)])
        b = bb.build()
        H = int(np.prod(nopts))
        got = np.zeros(H * len(nopts), np.int32)
        n = ref.ref_enumerate_haplotypes(C.byref(b), 0, ptr(got, c_i32p))
        assert n == H
        arr = np.array(nopts, np.int32)
        for h in range(H):
            mine = np.zeros(len(nopts), np.int32)
            o.oracle_hap_options(len(nopts), ptr(arr, c_i32p), h, ptr(mine, c_i32p))
            assert list(mine) == list(got[h * len(nopts):(h + 1) * len(nopts)]), (nopts, h)


@needs_ref
@pytest.mark.parametrize("name", [n for n, _ in cases.SYNTH_CASES])
def test_align_synthetic_bit_exact(name):
    s = cases.synth(name)
    a = checkers.align(checkers.oracle(), "oracle_", s.batch, s.n_out)
    b = checkers.align(ref, "ref_", s.batch, s.n_out)
    assert np.array_equal(a, b), "max |diff| %g" % np.abs(a - b).max()


HANDMADE = [
    dict(seed=1), dict(seed=2, motif="ACG", copies=6), dict(seed=3, motif="A", copies=12, rep_opts=4),
    dict(seed=4, flank_opts=(2, 1)), dict(seed=5, flank_opts=(2, 3), rep_opts=2),
    dict(seed=6, homopolymer_edges=True, flank_opts=(2, 2), rep_opts=3, motif="A", copies=9),
    dict(seed=7, homopolymer_edges=True, rep_opts=4, motif="AT", copies=7),
    dict(seed=8, motif="AGAT", copies=3, rep_opts=5),       # alleles shorter than 6 repeat units
    dict(seed=9, motif="ACGTAC", copies=1, rep_opts=2),     # single copy: no deletion artefact possible
    dict(seed=10, qual_lo=-5, qual_hi=60),                  # qualities outside '!'..'J' are clamped
]


@needs_ref
@pytest.mark.parametrize("kw", HANDMADE, ids=lambda k: "-".join("%s=%s" % i for i in k.items()))
def test_align_handmade_bit_exact(kw):
    b = cases.handmade_batch(**kw)
    a, pa = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, want_pos=True)
    r = checkers.align(ref, "ref_", b, b.n_out)
    assert np.array_equal(a, r), "max |diff| %g" % np.abs(a - r).max()


@needs_ref
@pytest.mark.parametrize("kw", [dict(seed=21, flank_opts=(2, 2), rep_opts=3, homopolymer_edges=True, motif="A", copies=9),
                                dict(seed=22, rep_opts=6), dict(seed=23, flank_opts=(3, 1), rep_opts=2)],
                         ids=["homop", "rep6", "flank3"])
def test_align_masks_bit_exact_and_untouched(kw):
    blocks, reads = cases.handmade(**kw)
    H = cases.n_haps_of(blocks)
    rng = np.random.default_rng(kw["seed"])
    hap_mask = (rng.random(H) < 0.6).astype(np.uint8)
    hap_mask[rng.integers(0, H)] = 1
    pool_mask = (rng.random(len(reads)) < 0.7).astype(np.uint8)
    b = BatchBuilder().add_locus(blocks, reads).build(realign_pool=pool_mask, realign_hap=hap_mask)
    a = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, fill=123.25)
    r = checkers.align(ref, "ref_", b, b.n_out, fill=123.25)
    assert np.array_equal(a, r)
    a2 = a.reshape(len(reads), H)
    assert np.all(a2[pool_mask == 0] == 123.25)
    assert np.all(a2[:, hap_mask == 0] == 123.25)
    assert np.all(a2[pool_mask == 1][:, hap_mask == 1] != 123.25)


@needs_ref
def test_seedless_pool_gets_zero():
    blocks, reads = cases.handmade(seed=31)
    reads[3] = (reads[3][0], reads[3][1], -1)
    b = BatchBuilder().add_locus(blocks, reads).build()
    a = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, fill=5.0)
    r = checkers.align(ref, "ref_", b, b.n_out, fill=5.0)
    assert np.array_equal(a, r)
    H = cases.n_haps_of(blocks)
    assert np.all(a.reshape(-1, H)[3] == 0.0)


@needs_ref
def test_posteriors_bit_exact():
    for name in ("cfg1_plumbing", "cfg2_shape", "mates"):
        s = cases.synth(name)
        rng = np.random.default_rng(5)
        ll = -np.abs(rng.normal(40, 15, int(s.read_ll_size)))
        p1 = np.log(rng.uniform(0.05, 1.0, s.n_reads))
        p2 = np.log(rng.uniform(0.05, 1.0, s.n_reads))
        for haploid in (0, 1):
            hap = np.full(s.n_loci, haploid, np.uint8)
            args = (s.locus_read_off, s.locus_sample_off, s.n_haps, hap, ll, p1, p2, s.sample_label, s.read_weight)
            po, so, bo, to = checkers.posteriors(checkers.oracle(), "oracle_", *args)
            pr, sr, br, tr = checkers.posteriors(ref, "ref_", *args)
            assert np.array_equal(po, pr) and np.array_equal(so, sr) and np.array_equal(bo, br) and np.array_equal(to, tr)


@needs_ref
def test_seed_selection_matches_reference():
    rng = np.random.default_rng(9)
    o = checkers.oracle()
    from hipstr_b200.capi import load
    lib = load()
    n = 400
    starts, lens, coff, ctype, clen = [], [], [0], bytearray(), []
    for _ in range(n):
        st = int(rng.integers(900, 1010))
        ops, total = [], 0
        for _ in range(int(rng.integers(1, 7))):
            t = "=XID"[int(rng.choice(4, p=[0.6, 0.15, 0.12, 0.13]))]
            k = int(rng.integers(1, 60 if t == "=" else 6))
            ops.append((t, k))
        # make sure the first and last bases are not seeds of a 1-op read
        ops = [("X", 1)] + ops + [("X", 1)]
        for t, k in ops:
            ctype.extend(t.encode()); clen.append(k)
            if t != "D":
                total += k
        coff.append(len(clen)); starts.append(st); lens.append(total)
    starts, lens, coff, clen = (np.array(x, np.int32) for x in (starts, lens, coff, clen))
    rs, re_ = np.array([1000], np.int32), np.array([1048], np.int32)
    outs = []
    for f in (lib.hipstr_calc_seeds, o.oracle_calc_seeds, ref.ref_calc_seeds):
        out = np.zeros(n, np.int32)
        st = f(n, ptr(starts, c_i32p), ptr(lens, c_i32p), ptr(coff, c_i32p), bytes(ctype), ptr(clen, c_i32p), 960, 1088, 1,
               ptr(rs, c_i32p), ptr(re_, c_i32p), ptr(out, c_i32p))
        assert st == 0
        outs.append(out)
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[2])
    assert (outs[0] >= 0).sum() > 50 and (outs[0] < 0).sum() > 5
