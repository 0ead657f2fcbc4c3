"""Section 8(f) row 4: BAM access, read filtering / pairing, PCR-duplicate removal -- the product's host code against the
UNMODIFIED reference (bam_io.cpp over the vendored htslib, BamProcessor::read_and_filter_reads, remove_pcr_duplicates,
AdapterTrimmer, AlignmentFilters) driven through oracle/ref_bam_harness.cpp on real BAM files written by htslib."""
import ctypes as C
import os

import numpy as np
import pytest

import checkers
from hipstr_b200 import capi
from hipstr_b200.capi import c_f64p, c_i32p, ptr
from ingest_sim import ADAPTERS, Scenario, rand_seq, revcomp

needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def write_bams(sc, tmp_path):
    ref = checkers.ref()
    ref.ref_sam_to_bam.restype = C.c_int32
    ref.ref_sam_to_bam.argtypes = [C.c_char_p, C.c_char_p]
    paths = []
    for f in range(len(sc.files)):
        sam, bam = str(tmp_path / ("f%d.sam" % f)), str(tmp_path / ("f%d.bam" % f))
        with open(sam, "w") as fh:
            fh.write(sc.sam_text(f))
        assert ref.ref_sam_to_bam(sam.encode(), bam.encode()) == 0
        paths.append(bam)
    return paths


def ref_region_reads(paths, chrom, start, end):
    f = checkers.ref().ref_bam_region_reads
    f.restype = C.c_int32
    f.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    cap = 1 << 26
    buf = C.create_string_buffer(cap)
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    n = f(len(paths), arr, chrom.encode(), start, end, cap, buf)
    assert n >= 0
    return buf.raw[:n].decode("latin-1")


OPTION_ORDER = ("min_flank", "min_read_end_match", "maximal_end_match_window", "min_bp_before_indel", "require_paired_reads",
                "base_qual_trim", "max_total_reads", "max_mate_dist", "remove_pcr_dups", "trim_adapters")


def ref_filter(paths, sc, rg_map, opts):
    f = checkers.ref().ref_read_and_filter
    f.restype = C.c_int32
    cpp = C.POINTER(C.c_char_p)
    f.argtypes = [C.c_int32, cpp, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, cpp, cpp, cpp, c_i32p, C.c_double,
                  C.c_int32, C.c_char_p]
    mk = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])
    keys = list(rg_map)
    o = np.array([opts[k] for k in OPTION_ORDER], np.int32)
    cap = 1 << 26
    buf = C.create_string_buffer(cap)
    n = f(len(paths), mk(paths), b"chr1", sc.chrom.encode(), sc.region[0], sc.region[1], sc.period, len(keys), mk(keys),
          mk([rg_map[k][0] for k in keys]), mk([rg_map[k][1] for k in keys]), ptr(o, c_i32p), opts["min_sum_qual_log_prob"], cap, buf)
    assert n >= 0
    return buf.raw[:n].decode("latin-1")


DEFAULTS = dict(min_flank=5, min_read_end_match=10, maximal_end_match_window=15, min_bp_before_indel=7, require_paired_reads=1,
                base_qual_trim=ord("5"), max_total_reads=1000000, max_mate_dist=1000, remove_pcr_dups=1, trim_adapters=1,
                min_sum_qual_log_prob=-10.0)


def ours_filter(paths, sc, rg_map, opts):
    reader = capi.BamReader(paths)
    start = 0 if sc.region[0] < opts["max_mate_dist"] else sc.region[0] - opts["max_mate_dist"]
    recs = reader.fetch("chr1", start, sc.region[1] + opts["max_mate_dist"])
    filtered = recs.filter(sc.chrom, [sc.region], rg_map, **opts)
    return filtered.text(), filtered.counts(), filtered


# ---------------------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bam_region_queries_match_htslib(seed, tmp_path):
    sc = Scenario(seed, n_files=2, n_fragments=400)
    paths = write_bams(sc, tmp_path)
    reader = capi.BamReader(paths)
    rng = np.random.default_rng(seed)
    spans = [(3000, 5100), (0, 9000), (4000, 4001), (4050, 4050 + 1), (100, 200), (7000, 9000)]
    spans += [tuple(sorted(int(x) for x in rng.integers(2500, 5500, 2))) for _ in range(12)]
    total = 0
    for start, end in spans:
        if start == end:
            end += 1
        want = ref_region_reads(paths, "chr1", start, end)
        got = reader.fetch("chr1", start, end)
        assert got.text() == want, (start, end)
        total += len(got)
    assert total > 1000
    assert len(reader.fetch("chr2", 0, 5000)) == 0
    with pytest.raises(capi.HipstrError):
        reader.fetch("chrUn", 0, 10)
    groups = reader.read_groups()
    assert [(p, g) for p, g, _, _ in groups] == [(paths[f], g) for f in range(2) for g, _, _ in sc.files[f]["groups"]]
    assert [(s, l) for _, _, s, l in groups] == [(s, l) for f in range(2) for _, s, l in sc.files[f]["groups"]]


def test_bam_reader_reports_missing_files(tmp_path):
    with pytest.raises(capi.HipstrError):
        capi.BamReader([str(tmp_path / "absent.bam")])
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"this is not a BGZF file at all")
    with pytest.raises(capi.HipstrError):
        capi.BamReader([str(junk)])


@needs_ref
@pytest.mark.parametrize("seed,overrides", [
    (11, {}),
    (12, {}),
    (13, dict(require_paired_reads=0)),
    (14, dict(require_paired_reads=0, remove_pcr_dups=0)),
    (15, dict(base_qual_trim=ord(" "), trim_adapters=0)),
    (16, dict(min_flank=0, min_read_end_match=0, maximal_end_match_window=0, min_bp_before_indel=0)),
    (17, dict(max_mate_dist=300, min_sum_qual_log_prob=-25.0)),
    (18, dict(max_total_reads=20)),
    (19, dict(base_qual_trim=ord("?"), maximal_end_match_window=5, min_read_end_match=5, min_bp_before_indel=3, require_paired_reads=0)),
])
def test_read_and_filter_matches_reference(seed, overrides, tmp_path):
    sc = Scenario(seed, n_files=2, n_fragments=260)
    paths = write_bams(sc, tmp_path)
    rg_map = sc.rg_map(paths)
    opts = dict(DEFAULTS, **overrides)
    want = ref_filter(paths, sc, rg_map, opts)
    got, counts, _ = ours_filter(paths, sc, rg_map, opts)
    assert got == want
    kept = sum(1 for line in want.splitlines() if line[0] in "PU")
    assert counts["passed"] == kept
    if "max_total_reads" not in overrides:
        assert kept > 10 and counts["overlapping"] > kept


@needs_ref
def test_read_names_with_mate_suffixes(tmp_path):
    """name/1 and name/2 pair up (trim_alignment_name); PCR-duplicate removal is off because the reference asserts equal names."""
    sc = Scenario(31, n_files=1, n_fragments=200, name_suffix=True)
    paths = write_bams(sc, tmp_path)
    opts = dict(DEFAULTS, remove_pcr_dups=0)
    want = ref_filter(paths, sc, sc.rg_map(paths), opts)
    got, counts, _ = ours_filter(paths, sc, sc.rg_map(paths), opts)
    assert got == want and counts["passed"] > 10


@needs_ref
def test_filtered_view_feeds_snp_phasing(tmp_path):
    sc = Scenario(41, n_files=2, n_fragments=200)
    paths = write_bams(sc, tmp_path)
    text, counts, filtered = ours_filter(paths, sc, sc.rg_map(paths), DEFAULTS)
    v = filtered.view()
    lines = [l.split("\t") for l in text.splitlines()]
    names = [l[1] for l in lines if l[0] == "G"]
    assert [v.sample_names[i].decode() for i in range(v.n_samples)] == names
    b = v.reads
    n_entries = sum(1 for l in lines if l[0] in "PU")
    assert b.n_entries == n_entries == counts["passed"] and b.n_alns == sum(1 for l in lines if l[0] in "PMU")
    alns = [l for l in lines if l[0] in "PMU"]
    for a in range(b.n_alns):
        assert (b.aln_pos[a], b.aln_end[a]) == (int(alns[a][3]), int(alns[a][4]))
        assert C.string_at(b.bases + b.aln_seq_off[a], b.aln_seq_off[a + 1] - b.aln_seq_off[a]).decode() == alns[a][6]
        cigar = "".join("%d%s" % (b.cigar_len[c], chr(C.cast(b.cigar_type, C.POINTER(C.c_char))[c][0])) for c in range(b.aln_cigar_off[a], b.aln_cigar_off[a + 1]))
        assert cigar == alns[a][5]
    assert v.sample_entry_off[v.n_samples] == n_entries


# ---------------------------------------------------------------------------------------------------------------
def _one(fn, what, arg, arg2, flag, pos, end, bases, quals, ops):
    types = "".join(t for t, _ in ops).encode()
    lens = np.array([n for _, n in ops] + [0], np.int32)
    out_pos = np.zeros(3, np.int32)
    cap = len(bases) + 8
    seq, qual = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n_out = C.c_int32()
    ctype, clen = C.create_string_buffer(len(ops) + 8), np.zeros(len(ops) + 8, np.int32)
    rc = fn(what, arg, arg2, flag, pos, end, bases.encode(), quals.encode("latin-1"), len(ops), types, ptr(lens, c_i32p), ptr(out_pos, c_i32p), seq, qual,
            C.byref(n_out), ctype, ptr(clen, c_i32p))
    return rc, out_pos.tolist(), seq.value, qual.value, [(chr(ctype.raw[i]), int(clen[i])) for i in range(n_out.value)]


def _bind_trim(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_int32
    f.argtypes = [C.c_int32] * 6 + [C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, c_i32p, C.c_char_p, C.c_char_p, c_i32p, C.c_char_p, c_i32p]
    return f


def random_read(rng, chrom, with_adapter=False):
    sc = Scenario.__new__(Scenario)
    sc.rng, sc.chrom, sc.region, sc.period = rng, chrom, (len(chrom) // 2, len(chrom) // 2 + 30), 3
    pos = int(rng.integers(50, len(chrom) - 400))
    rev = bool(rng.random() < 0.5)
    seq, quals, ops = sc.make_read(pos, int(rng.integers(20, 151)), rev, int(rng.integers(1, 3)))
    ops = [o for o in ops if o[0] != "H"] if rng.random() < 0.7 else ops
    end = pos + sum(n for t, n in ops if t in "MD")
    return pos, end, seq, quals, ops, rev


@needs_ref
def test_trimming_steps_match_reference():
    rng = np.random.default_rng(5)
    chrom = rand_seq(rng, 3000)
    ours, ref = _bind_trim(capi.load(), "hipstr_trim_one"), _bind_trim(checkers.ref(), "ref_trim_one")
    changed = [0, 0, 0]
    for trial in range(1500):
        pos, end, seq, quals, ops, rev = random_read(rng, chrom)
        what = trial % 3
        flag = int(rng.choice([0, 16, 0x41, 0x51, 0x81, 0x91]))
        if what == 0:
            arg, arg2 = int(rng.choice([ord("5"), ord("#"), ord("?"), ord("I"), ord("~")])), 0
        elif what == 1:
            arg = arg2 = 0
        else:
            arg = int(rng.integers(0, len(seq) // 2 + 1))
            arg2 = int(rng.integers(0, len(seq) - arg + 1)) if rng.random() < 0.8 else 0
        want = _one(ref, what, arg, arg2, flag, pos, end, seq, quals, ops)
        got = _one(ours, what, arg, arg2, flag, pos, end, seq, quals, ops)
        assert got == want, (trial, what, arg, arg2, flag, pos, seq, quals, ops)
        changed[what] += want[2] != seq.encode()
    assert min(changed) > 20


@needs_ref
def test_adapter_trimming_hand_cases():
    """Exact adapter, one mismatch, overhang off the read end, too-short overlap, reverse-strand reads (5' trimming)."""
    ours, ref = _bind_trim(capi.load(), "hipstr_trim_one"), _bind_trim(checkers.ref(), "ref_trim_one")
    rng = np.random.default_rng(8)
    body = rand_seq(rng, 80)
    cases = []
    for name, ad in ADAPTERS.items():
        one_off = ad[:6] + ("A" if ad[6] != "A" else "C") + ad[7:]
        for tail in (ad, ad + "ACGTACGT", ad[:9], ad[:5], ad[:4], one_off, one_off[:9], ad[:3] + "TT" + ad[5:]):
            cases.append((body + tail, 0x41 if name != "r2" else 0x81))
            cases.append((revcomp(body + tail), 0x51 if name != "r2" else 0x91))
            cases.append((body + tail, 0))
            cases.append((revcomp(body + tail), 16))
    trimmed = 0
    for seq, flag in cases:
        ops = [("M", len(seq))]
        want = _one(ref, 1, 0, 0, flag, 500, 500 + len(seq), seq, "I" * len(seq), ops)
        got = _one(ours, 1, 0, 0, flag, 500, 500 + len(seq), seq, "I" * len(seq), ops)
        assert got == want, (seq, flag)
        trimmed += want[1][2] != len(seq)
    assert 20 < trimmed < len(cases)


def _filters(fn, pos, end, bases, quals, ops, chrom, window):
    fn.restype = C.c_int32
    fn.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, C.c_char_p, C.c_int32, c_i32p, c_f64p]
    types = "".join(t for t, _ in ops).encode()
    lens = np.array([n for _, n in ops] + [0], np.int32)
    out, s = np.zeros(5, np.int32), np.zeros(1)
    fn(pos, end, bases.encode(), quals.encode("latin-1"), len(ops), types, ptr(lens, c_i32p), chrom, window, ptr(out, c_i32p), ptr(s, c_f64p))
    return out.tolist(), float(s[0])


@needs_ref
def test_alignment_filters_match_reference():
    rng = np.random.default_rng(6)
    # a repetitive chromosome: the end-match filters only bite where the read could be shifted
    unit = rand_seq(rng, 7)
    chrom = rand_seq(rng, 600) + unit * 30 + rand_seq(rng, 300) + "AC" * 40 + rand_seq(rng, 600) + "T" * 30 + rand_seq(rng, 600)
    cchrom = chrom.encode()
    ref_fn, our_fn = checkers.ref().ref_alignment_filters, capi.load().hipstr_alignment_filters
    our_fn.restype = C.c_int32
    seen = set()
    for trial in range(1500):
        pos, end, seq, quals, ops, _ = random_read(rng, chrom)
        if rng.random() < 0.3:
            seq = seq.lower() if rng.random() < 0.5 else seq
        window = int(rng.choice([15, 5, 1, 40]))
        want = _filters(ref_fn, pos, end, seq, quals, ops, cchrom, window)
        got = _filters(our_fn, pos, end, seq, quals, ops, cchrom, window)
        assert got == want, (trial, pos, seq, ops, window)
        seen.add((want[0][0], want[0][3] >= 0, want[0][4] >= 0))
    assert len(seen) >= 4
    # reads hanging off the end of the reference sequence
    tail = chrom[-60:]
    for pos, ops in ((len(chrom) - 60, [("M", 60)]), (len(chrom) - 30, [("M", 60)]), (len(chrom) - 1, [("M", 5)]), (len(chrom) + 5, [("M", 5)])):
        seq = (tail + "ACGT" * 20)[:sum(n for _, n in ops)]
        args = (pos, pos + len(seq), seq, "I" * len(seq), ops, cchrom, 15)
        assert _filters(our_fn, *args) == _filters(ref_fn, *args), (pos, ops)


def test_golden_bam_fixture():
    """A committed BAM (tests/golden/make_bam_fixture.py, written by htslib through the reference harness) with the
    reference's own outputs next to it: the product is checked without the reference library being present."""
    bam = os.path.join(GOLDEN, "ingest_f0.bam")
    reader = capi.BamReader([bam])
    with open(os.path.join(GOLDEN, "ingest_f0.region.txt")) as fh:
        want_region = fh.read()
    assert reader.fetch("chr1", 3000, 5100).text().replace(bam, "BAM") == want_region
    import json
    with open(os.path.join(GOLDEN, "ingest_f0.meta.json")) as fh:
        meta = json.load(fh)
    recs = reader.fetch("chr1", meta["region"][0] - 1000, meta["region"][1] + 1000)
    rg_map = {bam + g: tuple(v) for g, v in meta["groups"].items()}
    filtered = recs.filter(meta["chrom"], [tuple(meta["region"])], rg_map)
    with open(os.path.join(GOLDEN, "ingest_f0.filtered.txt")) as fh:
        assert filtered.text() == fh.read()


@needs_ref
def test_bam_decoding_edge_cases(tmp_path):
    """Records htslib accepts that the simulator never writes: no sequence / no qualities, IUPAC codes (which the reference
    maps to blanks), every CIGAR operation, unmapped reads placed at their mate, all auxiliary value types, small integer types
    for AS / XS, a second chromosome."""
    head = ["@HD\tVN:1.5\tSO:coordinate", "@SQ\tSN:chr1\tLN:100000", "@SQ\tSN:chr2\tLN:5000", "@RG\tID:g\tSM:s\tLB:l", "@RG\tID:h\tSM:t"]
    recs = [
        ("plain", 99, "chr1", 1000, 60, "10M", "=", 1200, 210, "ACGTACGTAC", "IIIIIIIIII", ["RG:Z:g", "AS:i:10", "XS:i:3"]),
        ("noseq", 0, "chr1", 1001, 0, "5M", "*", 0, 0, "*", "*", ["RG:Z:g"]),
        ("noqual", 16, "chr1", 1002, 30, "4M", "*", 0, 0, "ACGT", "*", ["RG:Z:h", "XA:Z:chr2,+100,4M,0;"]),
        ("iupac", 0, "chr1", 1003, 30, "8M", "*", 0, 0, "ACRYKMNT", "!#5?IJ~~", ["RG:Z:g", "NM:i:2", "MD:Z:8"]),
        ("ops", 0, "chr1", 1004, 30, "2H3S4M2I3D5N2=1X1P2M4S", "*", 0, 0, "A" * 18, "I" * 18, ["RG:Z:g", "SA:Z:chr2,5,+,3S4M,60,0;"]),
        ("unmapped_mate", 69, "chr1", 1005, 0, "*", "=", 1005, 0, "ACGTAC", "IIIIII", ["RG:Z:g"]),
        ("tags", 0, "chr1", 1006, 30, "6M", "*", 0, 0, "ACGTAC", "IIIIII",
         ["XA:Z:chr1,-90,6M,1;chr2,+7,6M,0;", "AS:i:200", "XS:i:-3", "RG:Z:h", "ZF:f:1.5", "ZA:A:x", "ZB:B:c,1,-2,3", "ZS:B:S,1,2", "ZI:B:i,70000", "ZH:H:1AE3", "HP:i:2"]),
        ("big_as", 0, "chr1", 1007, 30, "6M", "*", 0, 0, "ACGTAC", "IIIIII", ["AS:i:70000", "XS:i:300", "RG:Z:g"]),
        ("long_" + "n" * 200, 147, "chr1", 1200, 60, "10M", "=", 1000, -210, "ACGTACGTAC", "IIIIIIIIII", ["RG:Z:g"]),
        ("far", 0, "chr1", 90000, 60, "10M", "*", 0, 0, "ACGTACGTAC", "IIIIIIIIII", ["RG:Z:g"]),
        ("other", 0, "chr2", 100, 60, "10M", "chr1", 1000, 0, "ACGTACGTAC", "IIIIIIIIII", ["RG:Z:g"]),
    ]
    sam = tmp_path / "edge.sam"
    with open(sam, "w") as fh:
        fh.write("\n".join(head) + "\n")
        for r in recs:
            fh.write("\t".join([r[0], str(r[1]), r[2], str(r[3]), str(r[4]), r[5], r[6], str(r[7]), str(r[8]), r[9], r[10]] + r[11]) + "\n")
    bam = str(tmp_path / "edge.bam")
    ref = checkers.ref()
    ref.ref_sam_to_bam.restype = C.c_int32
    ref.ref_sam_to_bam.argtypes = [C.c_char_p, C.c_char_p]
    assert ref.ref_sam_to_bam(str(sam).encode(), bam.encode()) == 0
    reader = capi.BamReader([bam])
    for chrom, start, end in (("chr1", 0, 100000), ("chr1", 990, 1010), ("chr1", 1004, 1005), ("chr1", 1011, 1013), ("chr1", 1016, 1019),
                              ("chr1", 89990, 90001), ("chr2", 0, 5000), ("chr1", 50000, 60000)):
        assert reader.fetch(chrom, start, end).text() == ref_region_reads([bam], chrom, start, end), (chrom, start, end)
    assert len(reader.fetch("chr1", 0, 100000)) == 10
    assert reader.read_groups() == [(bam, "g", "s", "l"), (bam, "h", "t", None)]
