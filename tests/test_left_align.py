"""left alignment of the raw reads (TrimAlignment, convertAlignment, realign, sequence-level reuse).

Raw BAM-like alignments are derived from the synthetic (already left-aligned) reads: = / X become M, the STR indel is moved
to another equivalent place inside the repeat, some reads get soft clips.  CPU: every read through the product's host
steps (with the NW operation string from the oracle, itself pinned to the reference in test_nw.py) must equal the
UNMODIFIED reference's TrimAlignment + convertAlignment / realign on an in-memory BamAlignment.  GPU: the batched call
(K6 + the reference's reuse-by-sequence loop) must equal that loop replayed in Python over the reference's per-read results."""
import ctypes as C

import numpy as np
import pytest

import checkers
from hipstr_b200.capi import Synth, c_i32p, load, make_locus_reads, ptr
from test_nw import _bind as bind_nw, _call as call_nw

needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")


def raw_reads(s, seed=0, clip_rate=0.15, lower_rate=0.1):
    """[(pos, end_pos exclusive, bases, quals, cigar)] per read of the Synth, plus per-locus chromosome and trim bounds."""
    from ref_genotyper import LocusReads
    rng = np.random.default_rng(seed)
    reads, chroms, lro = [], [], [0]
    period = int(s.cfg.period) or 4
    for l in range(s.n_loci):
        rd = LocusReads(s, l)
        chroms.append(rd.chrom_seq)
        for r in range(rd.n_reads):
            b = bytes(rd.bases[rd.seq_off[r]:rd.seq_off[r + 1]]).decode()
            q = bytes(rd.quals[rd.seq_off[r]:rd.seq_off[r + 1]]).decode()
            cig = [(chr(rd.cigar_type[c]), int(rd.cigar_len[c])) for c in range(rd.cigar_off[r], rd.cigar_off[r + 1])]
            # = / X -> M, merged
            ops = []
            for t, n in cig:
                t = "M" if t in "=X" else t
                if ops and ops[-1][0] == t:
                    ops[-1] = (t, ops[-1][1] + n)
                else:
                    ops.append((t, n))
            # move the indel k motif copies to the right (an equally valid placement inside the repeat)
            for i, (t, n) in enumerate(ops):
                if t in "ID" and 0 < i < len(ops) - 1 and ops[i + 1][0] == "M":
                    k = int(rng.integers(0, 4)) * period
                    k = min(k, ops[i + 1][1] - 1)
                    if k > 0:
                        ops = ops[:i] + [("M", k), (t, n), ("M", ops[i + 1][1] - k)] + ops[i + 2:]
                        if ops[i - 1][0] == "M":
                            ops = ops[:i - 1] + [("M", ops[i - 1][1] + k)] + ops[i + 1:]
                    break
            pos = int(rd.start[r])
            end_pos = pos + sum(n for t, n in ops if t in "MD")
            if rng.random() < clip_rate and ops[0][0] == "M" and ops[0][1] > 8:   # soft-clip the first bases
                k = int(rng.integers(1, 6))
                ops = [("S", k), ("M", ops[0][1] - k)] + ops[1:]
                pos += k
            if rng.random() < lower_rate:
                b = b.lower()
            reads.append((pos, end_pos, b, q, ops))
        lro.append(len(reads))
    start, stop = int(s.view.region_start), int(s.view.region_stop)
    return reads, chroms, np.array(lro, np.int32), (start - 40 if start > 40 else 1), stop + 40


def ref_one(lib, read, chrom, trim):
    pos, end_pos, b, q, ops = read
    n = len(b) + 8
    out_pos = np.zeros(2, np.int32)
    seq, qual, aln = (C.create_string_buffer(2 * n) for _ in range(3))
    ncig = C.c_int32()
    ctype = C.create_string_buffer(2 * n)
    clen = np.zeros(2 * n, np.int32)
    t = "".join(o[0] for o in ops).encode()
    ln = np.array([o[1] for o in ops], np.int32)
    how = lib.ref_left_align_one(pos, end_pos, b.encode(), q.encode(), len(ops), t, ptr(ln, c_i32p), chrom, 1 if trim else 0,
                                 trim[0] if trim else 0, trim[1] if trim else 0, ptr(out_pos, c_i32p), seq, qual, aln, C.byref(ncig),
                                 ctype, ptr(clen, c_i32p))
    if how < 0:
        return how, None
    return how, (int(out_pos[0]), int(out_pos[1]), seq.value.decode(), qual.value.decode(),
                 [(ctype.raw[i:i + 1].decode(), int(clen[i])) for i in range(ncig.value)])


def ours_one(lib, nw, read, chrom, trim):
    pos, end_pos, b, q, ops = read
    n = len(b) + 8
    out_pos, window = np.zeros(2, np.int32), np.zeros(2, np.int32)
    seq, qual = C.create_string_buffer(2 * n), C.create_string_buffer(2 * n)
    ncig = C.c_int32()
    ctype = C.create_string_buffer(2 * n)
    clen = np.zeros(2 * n, np.int32)
    t = "".join(o[0] for o in ops).encode()
    ln = np.array([o[1] for o in ops], np.int32)
    args = [pos, end_pos, b.encode(), q.encode(), len(ops), t, ptr(ln, c_i32p), chrom, 1 if trim else 0, trim[0] if trim else 0,
            trim[1] if trim else 0]
    tail = [ptr(window, c_i32p), ptr(out_pos, c_i32p), seq, qual, C.byref(ncig), ctype, ptr(clen, c_i32p)]
    how = lib.hipstr_left_align_one(*args, None, *tail)
    if how == 3:   # needs the alignment of the trimmed read against its window
        win = chrom[window[0]:window[0] + window[1]].decode()
        ops_str, _ = call_nw(nw, win, seq.value.decode(), False)
        how = lib.hipstr_left_align_one(*args, ops_str.encode(), *tail)
    if how < 0:
        return how, None
    return how, (int(out_pos[0]), int(out_pos[1]), seq.value.decode(), qual.value.decode(),
                 [(ctype.raw[i:i + 1].decode(), int(clen[i])) for i in range(ncig.value)])


def _bind_ref(lib):
    lib.ref_left_align_one.restype = C.c_int32
    lib.ref_left_align_one.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, C.c_char_p,
                                       C.c_int32, C.c_int32, C.c_int32, c_i32p, C.c_char_p, C.c_char_p, C.c_char_p, c_i32p,
                                       C.c_char_p, c_i32p]
    return lib


CASES = [dict(n_loci=2, n_samples=6, reads_per_sample=10, n_alleles=6, read_len=150, seed=301, stutter_rate=0.2),
         dict(n_loci=2, n_samples=5, reads_per_sample=10, n_alleles=5, read_len=110, seed=302, period=2, ref_copies=15, stutter_rate=0.3,
              sub_rate=0.02),
         dict(n_loci=2, n_samples=5, reads_per_sample=8, n_alleles=6, read_len=250, seed=303, trim=0),
         dict(n_loci=2, n_samples=5, reads_per_sample=10, n_alleles=5, read_len=120, seed=304, period=1, ref_copies=14, stutter_rate=0.3)]


@needs_ref
@pytest.mark.parametrize("kw", CASES, ids=lambda k: "seed%d" % k["seed"])
@pytest.mark.parametrize("trim", [True, False])
def test_host_steps_match_reference_read_by_read(kw, trim):
    s = Synth(**kw)
    reads, chroms, lro, t0, t1 = raw_reads(s)
    ref, ours, nw = _bind_ref(checkers.ref()), load(), bind_nw(checkers.oracle(), "oracle_nw_align")
    n_realigned = n_converted = 0
    for l in range(s.n_loci):
        for r in range(lro[l], lro[l + 1]):
            want = ref_one(ref, reads[r], chroms[l], (t0, t1) if trim else None)
            got = ours_one(ours, nw, reads[r], chroms[l], (t0, t1) if trim else None)
            assert got == want, (r, reads[r])
            n_realigned += want[0] == 2
            n_converted += want[0] == 1
    assert n_realigned > 20 and n_converted > 5


def python_loop(per_read, trimmed):
    """GenotyperBamProcessor::left_align_reads' reuse-by-sequence loop (:59-79) over the reference's per-read results;
    trimmed[r] = (bases, qualities) of read r after TrimAlignment, the key of the reuse map."""
    out, seen = [], {}
    for r, (how, aln) in enumerate(per_read):
        if how == -1:
            continue
        key, quals = trimmed[r]
        prev = seen.get(key)
        if prev is not None and len(prev[2]) == len(key):
            out.append((r, (prev[0], prev[1], key.upper(), quals, prev[4])))
            continue
        if how == 0:
            continue
        seen[key] = aln
        out.append((r, aln))
    return out


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES[:3], ids=lambda k: "seed%d" % k["seed"])
def test_batched_left_alignment_matches_reference_loop(kw):
    from hipstr_b200.capi import Context, LeftAligned
    s = Synth(**kw)
    reads, chroms, lro, t0, t1 = raw_reads(s)
    ref = _bind_ref(checkers.ref())
    R = len(reads)
    raw = make_locus_reads(lro, s.locus_sample_off, reads, s.sample_label, np.arange(R), s.log_p1, s.log_p2, s.haploid)
    ctx = Context(0)
    la = LeftAligned(ctx, s.n_loci, raw, chroms, [t0] * s.n_loci, [t1] * s.n_loci)
    got, got_lro = la.reads()
    want = []
    for l in range(s.n_loci):
        locus_raw = reads[lro[l]:lro[l + 1]]
        per_read = [ref_one(ref, rd, chroms[l], (t0, t1)) for rd in locus_raw]
        trimmed = [trim_like_reference(rd, t0, t1) for rd in locus_raw]   # the reuse key keeps the read's original case
        for i, aln in python_loop(per_read, trimmed):
            want.append((lro[l] + i, aln))
    assert [int(x) for x in la.source] == [w[0] for w in want]
    assert got == [w[1] for w in want]
    assert la.nw_alignments > 0 and la.failed == 0
    # left-aligning the displaced indels recovers the generator's own left-aligned CIGARs for most reads
    from ref_genotyper import LocusReads
    same = total = 0
    for l in range(s.n_loci):
        rd = LocusReads(s, l)
        for r in range(rd.n_reads):
            cig = [(chr(rd.cigar_type[c]), int(rd.cigar_len[c])) for c in range(rd.cigar_off[r], rd.cigar_off[r + 1])]
            idx = np.nonzero(la.source == lro[l] + r)[0]
            if len(idx):
                total += 1
                same += got[idx[0]][4] == cig
    if kw.get("trim", 1):   # (untrimmed generator reads are cut to +-40 bp here, so their CIGARs cannot be compared)
        assert same > 0.6 * total, (same, total)
    la.close()
    ctx.close()


@needs_ref
@pytest.mark.gpu
def test_raw_reads_to_vcf_through_left_alignment():
    """The chain a caller runs: raw BAM-level reads -> hipstr_left_align_reads_host (K6) -> hipstr_genotyper_create_from_reads ->
    genotype() -> write_vcf; the reference SeqStutterGenotyper gets the same left-aligned reads."""
    from hipstr_b200.capi import Context, Genotyper, LeftAligned
    from ref_genotyper import ReadsOfLocus, RefGenotyper
    kw = dict(n_loci=3, n_samples=8, reads_per_sample=15, n_alleles=5, read_len=120, seed=305, stutter_rate=0.2)
    s = Synth(**kw)
    reads, chroms, lro, t0, t1 = raw_reads(s, seed=3)
    R = len(reads)
    raw = make_locus_reads(lro, s.locus_sample_off, reads, s.sample_label, np.arange(R), s.log_p1, s.log_p2, s.haploid)
    ctx = Context(0)
    la = LeftAligned(ctx, s.n_loci, raw, chroms, [t0] * s.n_loci, [t1] * s.n_loci)
    aligned, alro = la.reads()
    L = s.n_loci
    start, stop, period = int(s.view.region_start), int(s.view.region_stop), int(s.cfg.period) or 4
    g = Genotyper.from_reads(ctx, la.view, L, [start] * L, [stop] * L, [period] * L, chroms)
    ok = g.genotype(1000, 4, 0.01, True)
    names = ["S%d" % i for i in range(kw["n_samples"])]
    loci = g.vcf_loci(["chrS"] * L, ["STR"] * L, [start] * L, [stop] * L, [period] * L, chroms, names * L, names)
    records = g.write_vcf(loci)
    for l in range(L):
        src = la.source[alro[l]:alro[l + 1]]
        rd = ReadsOfLocus(aligned[alro[l]:alro[l + 1]], kw["n_samples"], s.sample_label[src], src, s.log_p1[src], s.log_p2[src], chroms[l],
                          (start, stop), period)
        r = RefGenotyper(rd, reassemble_flanks=True)
        assert r.initialized and r.genotype() == bool(ok[l])
        assert g.blocks(l) == [b[3] for b in r.blocks()]
        assert np.array_equal(g.results(l)["best"], r.results()["best"])
        assert records[l][1].replace(":-0.00:", ":0.00:") == r.vcf().rstrip("\n").replace(":-0.00:", ":0.00:")
    g.close()
    la.close()
    ctx.close()


def trim_like_reference(read, t0, t1):
    """Bases / qualities left after BamAlignment::TrimAlignment (quality bound '~' never stops the trimming)."""
    pos, end_pos, b, q, ops = read
    ops = [list(o) for o in ops]
    lt = rt = 0
    sp, ep = pos, end_pos
    while sp < t0 and ops:
        t = ops[0][0]
        if t in "M=X":
            lt += 1
            sp += 1
        elif t == "D":
            sp += 1
        elif t in "IS":
            lt += 1
        ops[0][1] -= 1
        if ops[0][1] == 0:
            ops.pop(0)
    while ep > t1 and ops:
        t = ops[-1][0]
        if t in "M=X":
            rt += 1
            ep -= 1
        elif t == "D":
            ep -= 1
        elif t in "IS":
            rt += 1
        ops[-1][1] -= 1
        if ops[-1][1] == 0:
            ops.pop()
    return b[lt:len(b) - rt], q[lt:len(q) - rt]
