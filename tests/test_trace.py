"""a16: alignment traceback (HapAligner::trace_optimal_aln / retrace).
CPU: the oracle restatement reproduces the compiled reference exactly -- alignment-operation strings, stutter
sizes, flank indel / SNP lists and, through read[span], the STR / flank sequences the reference keeps as strings.
GPU (K5): every output must be identical to the oracle's (all integers / bytes)."""
import ctypes as C

import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import MAX_BLOCKS, AlignBatch, BatchBuilder, TraceOut, c_i32p, load, trace_batch

KEYS = ("stutter_size", "span_start", "span_len", "flank_ins", "flank_del", "n_indels", "indels", "n_snps", "snps")


def _fn(lib, name, extra=False):
    f = getattr(lib, name)
    f.restype = C.c_int32
    f.argtypes = [C.POINTER(AlignBatch), c_i32p, C.c_int32, c_i32p, c_i32p, C.POINTER(TraceOut)] + ([C.c_char_p, C.c_int32] if extra else [])
    return f


def block_starts(batch, first=1000):
    """Genomic starts: blocks of a locus tile the reference from `first` on (reference allele lengths)."""
    lbo = np.ctypeslib.as_array(batch.locus_block_off, shape=(batch.n_loci + 1,))
    boo = np.ctypeslib.as_array(batch.block_opt_off, shape=(batch.n_blocks + 1,))
    oso = np.ctypeslib.as_array(batch.opt_seq_off, shape=(batch.n_options + 1,))
    out = []
    for l in range(batch.n_loci):
        pos = first
        for b in range(lbo[l], lbo[l + 1]):
            out.append(pos)
            pos += int(oso[boo[b] + 1] - oso[boo[b]])
    return np.array(out, np.int32)


def synth_traces(name, n=120, seed=0):
    s = cases.synth(name)
    rng = np.random.default_rng(seed)
    pools = np.nonzero(s.pool_seed >= 0)[0][:n].astype(np.int32)
    loc = np.searchsorted(s.locus_pool_off, pools, side="right") - 1
    haps = np.array([rng.integers(0, s.n_haps[l]) for l in loc], np.int32)
    addr = C.c_void_p.from_buffer(s.batch, AlignBatch.pool_bases.offset).value
    pb = C.string_at(addr, int(s.pool_seq_off[-1]))
    reads = [pb[s.pool_seq_off[p]:s.pool_seq_off[p + 1]].decode() for p in pools]
    return s, s.batch, pools, haps, reads


def hand_traces(kw, seed=0):
    blocks, reads = cases.handmade(**kw)
    b = BatchBuilder().add_locus(blocks, reads).build()
    rng = np.random.default_rng(seed)
    H = cases.n_haps_of(blocks)
    pools = np.repeat(np.arange(len(reads)), 2).astype(np.int32)
    haps = rng.integers(0, H, len(pools)).astype(np.int32)
    return None, b, pools, haps, [reads[p][0] for p in pools]


SYNTH = ["cfg1_plumbing", "cfg2_shape", "period2", "period1_homopolymer", "period3_noisy", "short_reads", "long_untrimmed"]
HAND = [dict(seed=1), dict(seed=4, flank_opts=(2, 1)), dict(seed=6, homopolymer_edges=True, flank_opts=(2, 2), rep_opts=3, motif="A", copies=9),
        dict(seed=8, motif="AGAT", copies=3, rep_opts=5), dict(seed=10, qual_lo=-5, qual_hi=60),
        dict(seed=12, n_reads=40, motif="AAAG", copies=20, rep_opts=3)]
ALL = [("synth", n) for n in SYNTH] + [("hand", k) for k in HAND]


def _load(case):
    kind, arg = case
    return synth_traces(arg) if kind == "synth" else hand_traces(arg)


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("case", ALL, ids=lambda c: str(c[1]))
def test_oracle_trace_equals_reference(case):
    keep, batch, pools, haps, reads = _load(case)
    bs = block_starts(batch)
    st, o = trace_batch(_fn(checkers.oracle(), "oracle_trace_batch"), batch, bs, pools, haps)
    assert st == 0
    stride = 1024
    buf = C.create_string_buffer(len(pools) * MAX_BLOCKS * stride)
    st, r = trace_batch(_fn(checkers.ref(), "ref_trace_batch", True), batch, bs, pools, haps, extra_args=(buf, C.c_int32(stride)))
    assert st == 0
    assert o["hap_aln"] == r["hap_aln"]
    for k in ("stutter_size", "flank_ins", "flank_del", "n_indels", "indels", "n_snps", "snps"):
        assert np.array_equal(o[k], r[k]), k
    for i in range(len(pools)):       # read[span] reproduces the reference's str_seq / flank_seq strings
        for b in range(MAX_BLOCKS):
            want = buf.raw[(i * MAX_BLOCKS + b) * stride:(i * MAX_BLOCKS + b + 1) * stride].split(b"\0")[0].decode()
            assert reads[i][o["span_start"][i][b]:o["span_start"][i][b] + o["span_len"][i][b]] == want, (i, b)
    # the operation string consumes exactly the read: M + I + S == read length
    for a, rd in zip(o["hap_aln"], reads):
        assert sum(a.count(c) for c in "MIS") == len(rd)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ALL, ids=lambda c: str(c[1]))
def test_gpu_trace_equals_oracle(case):
    from hipstr_b200.capi import Context
    keep, batch, pools, haps, reads = _load(case)
    bs = block_starts(batch)
    st, o = trace_batch(_fn(checkers.oracle(), "oracle_trace_batch"), batch, bs, pools, haps)
    assert st == 0
    ctx = Context(0)
    g = ctx.trace(batch, bs, pools, haps)
    ctx.close()
    bad = [i for i in range(len(pools)) if g["hap_aln"][i] != o["hap_aln"][i]]
    if bad:
        i = bad[0]
        print("first mismatch trace %d pool %d hap %d\n gpu    %s\n oracle %s" % (i, pools[i], haps[i], g["hap_aln"][i], o["hap_aln"][i]))
    assert not bad, "%d of %d alignment strings differ" % (len(bad), len(pools))
    assert np.array_equal(g["seed_hap_pos"], o["seed_hap_pos"])
    for k in KEYS:
        assert np.array_equal(g[k], o[k]), k


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("case", ALL, ids=lambda c: str(c[1]))
def test_stitch_trace_matches_reference(case):
    """hipstr_stitch_trace (host) on the oracle's trace reproduces AlignmentTrace::traced_aln() of the reference:
    start, stop, CIGAR and gapped alignment, given the reference's own haplotype-vs-reference alignment string."""
    from hipstr_b200.capi import stitch_trace
    keep, batch, pools, haps, reads = _load(case)
    bs = block_starts(batch)
    st, o = trace_batch(_fn(checkers.oracle(), "oracle_trace_batch"), batch, bs, pools, haps)
    assert st == 0
    ref = checkers.ref()
    f = ref.ref_trace_stitched
    f.restype = C.c_int32
    f.argtypes = [C.POINTER(AlignBatch), c_i32p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, c_i32p, c_i32p,
                  C.c_char_p, C.c_char_p]
    lbo = np.ctypeslib.as_array(batch.locus_block_off, shape=(batch.n_loci + 1,))
    lpo = np.ctypeslib.as_array(batch.locus_pool_off, shape=(batch.n_loci + 1,))
    seeds = np.ctypeslib.as_array(batch.pool_seed, shape=(batch.n_pools,))
    n_checked = 0
    for i in range(0, len(pools), 3):
        cap = 4096
        b1, b2, b3, b4 = (C.create_string_buffer(cap) for _ in range(4))
        a, b = C.c_int32(), C.c_int32()
        st = f(C.byref(batch), bs.ctypes.data_as(c_i32p), int(pools[i]), int(haps[i]), b1, b2, cap, C.byref(a), C.byref(b), b3, b4)
        assert st == 0
        assert b2.value.decode() == o["hap_aln"][i]
        locus = int(np.searchsorted(lpo, pools[i], side="right") - 1)
        hap_start = int(bs[lbo[locus]])
        st, start, stop, cigar, aln = stitch_trace(hap_start, b1.value.decode(), o["hap_aln"][i], int(o["seed_hap_pos"][i]),
                                                   int(seeds[pools[i]]), reads[i])
        assert st == 0
        assert (start, stop, cigar, aln) == (a.value, b.value, b3.value.decode(), b4.value.decode()), i
        # the span-only form (no CIGAR / alignment buffers) the loop uses
        lib = load()
        s2, e2, n2 = C.c_int32(), C.c_int32(), C.c_int32(-1)
        st = lib.hipstr_stitch_trace(hap_start, b1.value, o["hap_aln"][i].encode(), int(o["seed_hap_pos"][i]), int(seeds[pools[i]]),
                                     reads[i].encode(), C.byref(s2), C.byref(e2), 0, None, None, C.byref(n2), 0, None)
        assert (st, s2.value, e2.value, n2.value) == (0, a.value, b.value, 0), i
        # ... and the indexed form (hipstr_hap_aln_index + hipstr_trace_span)
        index = np.zeros(3 * (len(b1.value) + 1), np.int32)
        assert lib.hipstr_hap_aln_index(b1.value, len(b1.value), index.ctypes.data_as(c_i32p)) == 0
        s3, e3 = C.c_int32(), C.c_int32()
        st = lib.hipstr_trace_span(hap_start, b1.value, len(b1.value), index.ctypes.data_as(c_i32p), o["hap_aln"][i].encode(),
                                   int(o["seed_hap_pos"][i]), int(seeds[pools[i]]), C.byref(s3), C.byref(e3))
        assert (st, s3.value, e3.value) == (0, a.value, b.value), i
        n_checked += 1
    assert n_checked >= 10


def test_indexed_trace_span_equals_the_stepping_form():
    """hipstr_trace_span (jumps over the aligned part of the walk through an index of the haplotype's operation string) against
    hipstr_stitch_trace without string buffers (steps through both strings) on random operation strings: haplotype strings
    over M / I / D with runs of D, read strings over M / I / D / S with clipped and inserted ends, seeds anywhere (also beyond
    either string), an occasional foreign character.  Status, start and stop must agree in every case."""
    lib = load()
    rng = np.random.default_rng(12)
    n_ok = n_bad = 0
    for trial in range(6000):
        n_h = int(rng.integers(1, 60))
        hap = "".join(rng.choice(list("MMMMMMIDD"), n_h))
        if trial % 2 == 0:   # anything against anything: mostly inconsistent pairs
            n_r = int(rng.integers(1, 50))
            core = "".join(rng.choice(list("MMMMMMMIDD"), n_r))
            read = "".join(rng.choice(list("SI"), int(rng.integers(0, 4)))) + core + "".join(rng.choice(list("IS"), int(rng.integers(0, 4))))
            seed_hap_pos = int(rng.integers(-1, n_h + 2))
            seed_base = int(rng.integers(-1, len(read) + 2))
        else:                # a read laid over a stretch of the haplotype's bases, seeded on one of its matches
            n_bases = sum(c != "D" for c in hap)
            x0 = int(rng.integers(0, max(1, n_bases)))
            x1 = int(rng.integers(x0, n_bases + 1))
            ops = []
            for _ in range(x0, x1):
                if rng.random() < 0.06:
                    ops.append("I")
                ops.append("M" if rng.random() < 0.88 else "D")
            head = "".join(rng.choice(list("SI"), int(rng.integers(0, 3))))
            read = head + "".join(ops) + "".join(rng.choice(list("IS"), int(rng.integers(0, 3))))
            matches = [k for k, c in enumerate(ops) if c == "M"]
            if matches:
                j = int(rng.choice(matches))
                seed_hap_pos = x0 + sum(c != "I" for c in ops[:j])
                seed_base = len(head) + sum(c != "D" for c in ops[:j])
            else:
                seed_hap_pos, seed_base = x0, len(head)
        if trial % 97 == 0 and read:
            k = int(rng.integers(0, len(read)))
            read = read[:k] + "X" + read[k + 1:]
        if not read:
            read = "S"
        hap_start = int(rng.integers(0, 1000))
        a, b, n = C.c_int32(-7), C.c_int32(-7), C.c_int32(-1)
        st1 = lib.hipstr_stitch_trace(hap_start, hap.encode(), read.encode(), seed_hap_pos, seed_base, b"", C.byref(a), C.byref(b), 0,
                                      None, None, C.byref(n), 0, None)
        index = np.zeros(3 * (n_h + 1), np.int32)
        assert lib.hipstr_hap_aln_index(hap.encode(), n_h, index.ctypes.data_as(c_i32p)) == 0
        c, d = C.c_int32(-7), C.c_int32(-7)
        st2 = lib.hipstr_trace_span(hap_start, hap.encode(), n_h, index.ctypes.data_as(c_i32p), read.encode(), seed_hap_pos, seed_base,
                                    C.byref(c), C.byref(d))
        assert st1 == st2, (trial, hap, read, seed_hap_pos, seed_base, st1, st2)
        if st1 == 0:
            assert (a.value, b.value) == (c.value, d.value), (trial, hap, read, seed_hap_pos, seed_base)
            n_ok += 1
        else:
            n_bad += 1
    assert n_ok > 1500 and n_bad > 300, (n_ok, n_bad)
    index = np.zeros(12, np.int32)
    assert lib.hipstr_hap_aln_index(b"MXM", 3, index.ctypes.data_as(c_i32p)) != 0   # only M / I / D can be indexed


# ---- complete flank lists (hipstr_trace_flank_lists): fixed-size slots of K5 vs reads with more entries ----------
def noisy_flank_traces(seed=5):
    """Reads with dozens of confident flank mismatches and a few flank indels over a 330-bp flank (what a chimeric or
    mismapped read looks like).  The model lets a read hang off the haplotype ends for free, so the best path clips such a
    read instead of paying for its mismatches: the fixed slots of K5 (16 indels, 32 SNPs) are not reachable through a
    maximum-likelihood path on real flank lengths -- the case checks that the lists stay identical to the reference's."""
    rng = np.random.default_rng(seed)
    left = "".join("ACGT"[i] for i in rng.integers(0, 4, 330))
    right = "".join("ACGT"[i] for i in rng.integers(0, 4, 60))
    blocks = [(0, [left]), (2, ["AC" * 8, "AC" * 9]), (0, [right])]
    hap = left + "AC" * 8 + right
    reads = []
    for k in range(6):
        rd = list(hap)
        for i in range(4 + k, 320, 9):
            rd[i] = "ACGT"[("ACGT".index(rd[i]) + 1 + k % 3) % 4]
        if k >= 3:                       # and a few flank indels
            del rd[len(hap) - 20]
            rd.insert(100, "G")
        rd = "".join(rd)
        reads.append((rd, "9" * len(rd), len(rd) - 25))    # seed in the right flank: the left side spans the repeat and the long flank
    b = BatchBuilder().add_locus(blocks, reads).build()
    pools = np.arange(len(reads), dtype=np.int32)
    haps = np.array([0, 1, 0, 1, 0, 1], np.int32)
    return b, pools, haps


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("case", ALL[:6] + [("noisy", None)], ids=lambda c: str(c[1]))
def test_flank_lists_match_reference(case):
    """hipstr_trace_flank_lists rebuilds the reference's complete flank_indel_data / flank_snp_data from a trace's operation
    string -- also for reads with more entries than the fixed slots hold."""
    from hipstr_b200 import capi
    from hipstr_b200.capi import MAX_TRACE_SNPS, trace_flank_lists
    if case[0] == "noisy":
        batch, pools, haps = noisy_flank_traces()
    else:
        keep, batch, pools, haps, _ = _load(case)   # `keep` owns the batch's memory
    bs = block_starts(batch)
    st, o = trace_batch(_fn(checkers.oracle(), "oracle_trace_batch"), batch, bs, pools, haps, aln_stride=2048)
    assert st == 0
    lib = capi.load()
    for i in range(len(pools)):
        ind, snp = trace_flank_lists(lib, batch, bs, pools[i], haps[i], o["hap_aln"][i], o["seed_hap_pos"][i], o["stutter_size"][i])
        rind, rsnp = trace_flank_lists(checkers.ref(), batch, bs, pools[i], haps[i], None, None, None, name="ref_trace_lists")
        assert np.array_equal(ind, rind) and np.array_equal(snp, rsnp), i
        assert len(ind) == o["n_indels"][i] and len(snp) == o["n_snps"][i]
        # a caller with too few slots gets the TRUE counts and the first entries (the contract of the fixed-size K5 outputs)
        if len(snp) + len(ind) > 0:
            ind1, snp1 = trace_flank_lists(lib, batch, bs, pools[i], haps[i], o["hap_aln"][i], o["seed_hap_pos"][i], o["stutter_size"][i], cap=1)
            assert np.array_equal(ind1, ind) and np.array_equal(snp1, snp)


@pytest.mark.gpu
def test_gpu_trace_counts_beyond_slots():
    """K5 on reads full of flank mismatches: counts, slots and the host rebuild of the complete lists agree."""
    from hipstr_b200.capi import Context, MAX_TRACE_SNPS, trace_flank_lists
    batch, pools, haps = noisy_flank_traces()
    bs = block_starts(batch)
    st, o = trace_batch(_fn(checkers.oracle(), "oracle_trace_batch"), batch, bs, pools, haps, aln_stride=2048)
    ctx = Context(0)
    g = ctx.trace(batch, bs, pools, haps)
    assert g["hap_aln"] == o["hap_aln"]
    for k in KEYS:
        assert np.array_equal(g[k], o[k]), k
    for i in range(len(pools)):
        ind, snp = trace_flank_lists(ctx.lib, batch, bs, pools[i], haps[i], g["hap_aln"][i], g["seed_hap_pos"][i], g["stutter_size"][i])
        assert len(snp) == g["n_snps"][i] and np.array_equal(snp[:MAX_TRACE_SNPS], g["snps"][i][:min(len(snp), MAX_TRACE_SNPS)])
    ctx.close()
