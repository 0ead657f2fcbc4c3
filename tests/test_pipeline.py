"""BAM files -> VCF records: the chained stages (hipstr_b200/pipeline.py) against the UNMODIFIED reference program minus its
option parsing (GenotyperBamProcessor::process_regions over real BAM / FASTA / region files, oracle/ref_bam_harness.cpp
ref_process_regions).  Record lines must be identical (AB's "-0.00" normalised, see tests/test_vcf_record.py)."""
import ctypes as C

import numpy as np
import pytest

import checkers
from ingest_sim import MultiScenario
from test_ingest import write_bams

needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
canon = lambda t: t.replace(":-0.00:", ":0.00:")


def run_reference(paths, fasta, bed, out_vcf, def_stutter, min_total_reads=20, remove_dups=1, require_paired=1, recalc=0, gls=0, pls=0,
                  filters=0, snp_vcf=None, haploid=0, tenx=0, ref_vcf=None):
    f = checkers.ref().ref_process_regions
    f.restype = C.c_int32
    f.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.c_char_p, C.c_char_p]
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    o = np.array([def_stutter, min_total_reads, remove_dups, require_paired, recalc, gls, pls, filters, haploid, tenx], np.int32)
    assert f(len(paths), arr, fasta.encode(), bed.encode(), out_vcf.encode(), o.ctypes.data_as(C.POINTER(C.c_int32)),
             snp_vcf.encode() if snp_vcf else None, ref_vcf.encode() if ref_vcf else None) == 0
    import gzip
    with gzip.open(out_vcf, "rt") as fh:       # the reference always writes BGZF (bgzfostream)
        lines = fh.read().splitlines()
    header = [l for l in lines if l.startswith("#")]
    return header, [l for l in lines if not l.startswith("#")]


def files_of(sc, tmp_path, extra_regions=()):
    paths = write_bams(sc, tmp_path)
    fasta, bed = str(tmp_path / "ref.fa"), str(tmp_path / "regions.bed")
    text = sc.fasta_text()
    with open(fasta, "w") as fh:
        fh.write(text)
    with open(fasta + ".fai", "w") as fh:      # samtools faidx: name, length, offset of the first base, bases and bytes per line
        at = 0
        for block in text.split(">")[1:]:
            name, _, body = block.partition("\n")
            at += len(name) + 2
            n = len(body.replace("\n", ""))
            fh.write("%s\t%d\t%d\t60\t61\n" % (name, n, at))
            at += len(body)
    with open(bed, "w") as fh:
        fh.write(sc.region_text(extra_regions))
    return paths, fasta, bed


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("seed,def_stutter,kw", [
    (3, 1, {}),
    (7, 1, dict(snp_vcf=1)),                      # phasing log-likelihoods from a phased SNP VCF (K7)
    (8, 0, dict(snp_vcf=1, gls=1)),               # ... which also enter the EM stutter model
    (4, 0, {}),                                   # stutter models learned by the EM genotyper (K4)
    (5, 1, dict(require_paired=0, gls=1, pls=1, filters=1)),
    (6, 0, dict(recalc=1, remove_dups=0)),
    (9, 0, dict(haploid=1, snp_vcf=1)),           # --haploid-chrs chr1
    (10, 0, dict(tenx=1)),                        # --10x-bams: phasing from the HP tags (native driver only)
])
@pytest.mark.parametrize("driver", ["native", "staged"])
def test_bam_to_vcf_matches_reference(seed, def_stutter, kw, driver, tmp_path):
    from hipstr_b200 import capi, pipeline
    if kw.get("tenx") and driver == "staged":
        pytest.skip("the stage-by-stage Python chain does not read HP tags")
    sc = MultiScenario(seed, n_regions=4, n_fragments=220 if def_stutter else 600, hp_tags=bool(kw.get("tenx")))
    extra = [("chr1", 2000, 2200, 4, 50.0, "TOO_LONG"), ("chr1", 10, 40, 3, 10.0, "CONTIG_END"), ("chr1", 7000, 7030, 3, 10.0, "NO_READS")]
    paths, fasta, bed = files_of(sc, tmp_path, extra)
    kw = dict(kw)
    snp_vcf = write_snp_vcf(sc, tmp_path, seed)[1] if kw.pop("snp_vcf", 0) else None
    header, want = run_reference(paths, fasta, bed, str(tmp_path / "ref.vcf"), def_stutter, snp_vcf=snp_vcf, **kw)
    opt = pipeline.Options(min_total_reads=20, snp_vcf=snp_vcf, def_stutter_model=pipeline.DEFAULT_STUTTER if def_stutter else None,
                           recalc_stutter_model=bool(kw.get("recalc", 0)), haploid_chroms=("chr1",) if kw.get("haploid") else (), bams_from_10x=bool(kw.get("tenx")),
                           filter=dict(remove_pcr_dups=kw.get("remove_dups", 1), require_paired_reads=kw.get("require_paired", 1)))
    vcf_opt = dict(output_gls=kw.get("gls", 0), output_pls=kw.get("pls", 0), output_filters=kw.get("filters", 0))
    with capi.Context(0) as ctx:
        run = pipeline.process_regions if driver == "native" else pipeline.process_regions_staged
        records, summary = run(ctx, paths, pipeline.read_fasta(fasta), pipeline.read_regions(bed), opt, vcf_opt)
    print(summary)
    assert [canon(r[2]) for r in records] == [canon(w) for w in want]
    if snp_vcf:
        assert summary["phased_reads"] > 20
    assert len(want) >= 2 and summary["too_long"] == 1 and summary["near_contig_end"] == 1 and summary["too_few_reads"] >= 1
    # the sample columns of the header line are the sorted sample names the records were written for
    assert header[-1].split("\t")[9:] == sorted({s for f in sc.files for _, s, _ in f["groups"]})
    if driver == "native":
        assert summary["samples"] == header[-1].split("\t")[9:]
        # the whole FILE: header + records through hipstr::VCFWriter (BGZF), against the reference program's output file
        import gzip
        contigs = [("chr1", len(sc.chrom)), ("chr2", 5000), ("chr1_KI1_alt", 3000)]
        out_path = str(tmp_path / "ours.vcf.gz")
        pipeline.write_vcf_file(out_path, capi.vcf_header(fasta, "harness", contigs, summary["samples"], **vcf_opt), records)
        with gzip.open(out_path, "rt") as a, gzip.open(str(tmp_path / "ref.vcf"), "rt") as b:
            assert canon(a.read()) == canon(b.read())


def write_snp_vcf(sc, tmp_path, seed=1):
    ref = checkers.ref()
    ref.ref_vcf_bgzip_tabix.restype = C.c_int32
    ref.ref_vcf_bgzip_tabix.argtypes = [C.c_char_p, C.c_char_p]
    text, gz = str(tmp_path / "snps.vcf"), str(tmp_path / "snps.vcf.gz")
    with open(text, "w") as fh:
        fh.write(sc.snp_vcf_text(seed))
    assert ref.ref_vcf_bgzip_tabix(text.encode(), gz.encode()) == 0
    return text, gz


@needs_ref
@pytest.mark.parametrize("seed", [3, 4])
def test_snp_sets_match_create_snp_trees(seed, tmp_path):
    """hipstr_snp_vcf_region_sets against the reference's create_snp_trees over a bgzipped + tabix-indexed VCF (CPU only);
    the product reads the bgzipped file and the plain text alike."""
    from hipstr_b200.capi import SnpVcf
    sc = MultiScenario(seed, n_regions=3, n_fragments=10)
    text, gz = write_snp_vcf(sc, tmp_path, seed)
    f = checkers.ref().ref_snp_sets
    f.restype = C.c_int32
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    total = 0
    for path in (gz, text):
        vcf = SnpVcf(path)
        assert vcf.samples[0] == "X9" and vcf.has_chromosome("chr2") and not vcf.has_chromosome("chr3")
        for start, stop, period in sc.regions + [(60, 90, 3)]:
            for mate_dist, padding in ((1000, 15), (200, 0)):
                buf = C.create_string_buffer(1 << 22)
                n = f(gz.encode(), b"chr1", start, stop, period, mate_dist, padding, len(buf), buf)
                assert n >= 0
                got = vcf.region_sets("chr1", start - mate_dist if start > mate_dist else 1, stop + mate_dist, [(start, stop)], padding)
                off, pos, b1, b2 = got
                lines = []
                for s, name in enumerate(vcf.samples):
                    lines.append("S " + name)
                    lines += ["%d %s %s" % (pos[k], chr(b1[k]), chr(b2[k])) for k in range(off[s], off[s + 1])]
                assert "\n".join(lines) + "\n" == buf.raw[:n].decode(), (path, start, stop, mate_dist)
                total += len(pos)
        assert vcf.region_sets("chr3", 1, 1000) is None
    assert total > 60


@needs_ref
def test_snp_vcf_stream_errors(tmp_path):
    """The SNP VCF is streamed in bounded chunks: a BGZF file cut inside a block is an error (not a silently shorter panel),
    a file cut BETWEEN blocks parses, and a plain file without a final newline keeps its last record."""
    from hipstr_b200.capi import SnpVcf
    sc = MultiScenario(5, n_regions=2, n_fragments=8)
    text, gz = write_snp_vcf(sc, tmp_path, 5)
    raw = open(gz, "rb").read()
    cut = str(tmp_path / "cut.vcf.gz")
    with open(cut, "wb") as fh:
        fh.write(raw[:len(raw) // 2 if len(raw) > 200 else len(raw) - 10])
    with pytest.raises(Exception):
        SnpVcf(cut)
    whole = str(tmp_path / "no_eof_block.vcf.gz")     # the 28-byte BGZF end-of-file block removed: still whole gzip members
    with open(whole, "wb") as fh:
        fh.write(raw[:-28])
    full, part = SnpVcf(gz), SnpVcf(whole)
    assert part.samples == full.samples
    bare = str(tmp_path / "bare.vcf")
    with open(bare, "w") as fh:
        fh.write(open(text).read().rstrip("\n"))
    a, b = SnpVcf(text).region_sets("chr1", 1, 100000), SnpVcf(bare).region_sets("chr1", 1, 100000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[1]) > 0


@needs_ref
@pytest.mark.parametrize("flags", [dict(), dict(gls=1, pls=1, filters=1)])
def test_vcf_header_matches_reference(flags, tmp_path):
    """hipstr_vcf_header against the header the reference program writes (Genotyper::get_vcf_header through
    GenotyperBamProcessor::init_output_vcf); an empty region file, so no GPU is involved."""
    from hipstr_b200 import capi
    sc = MultiScenario(9, n_regions=1, n_fragments=5)
    paths, fasta, bed = files_of(sc, tmp_path)
    open(bed, "w").close()
    header, records = run_reference(paths, fasta, bed, str(tmp_path / "ref.vcf"), 1, **flags)
    assert not records
    contigs = [("chr1", len(sc.chrom)), ("chr2", 5000), ("chr1_KI1_alt", 3000)]
    samples = sorted({s for f in sc.files for _, s, _ in f["groups"]})
    got = capi.vcf_header(fasta, "harness", contigs, samples, output_gls=flags.get("gls", 0), output_pls=flags.get("pls", 0),
                          output_filters=flags.get("filters", 0))
    assert got == "\n".join(header) + "\n"


def test_process_regions_needs_a_device():
    """No CPU path: without a context the driver refuses (HIPSTR_ERR_NO_DEVICE), and bad arguments are rejected."""
    from hipstr_b200 import capi
    lib = capi.load()
    po, vo = capi.PipelineOptions(), capi.VcfOptions()
    lib.hipstr_pipeline_default_options(C.byref(po))
    lib.hipstr_vcf_default_options(C.byref(vo))
    assert (po.max_str_length, po.min_total_reads, po.filter.max_mate_dist, po.skip_padding) == (100, 100, 1000, 15)
    assert list(po.def_stutter_model) == [0.95, 0.05, 0.05, 0.95, 0.01, 0.01]
    h = C.c_void_p()
    mk = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])
    st = lib.hipstr_process_regions(None, 1, mk(["x.bam"]), None, 1, mk(["chr1"]), mk(["ACGT"]), 0, None, None, None, None, None, C.byref(po),
                                    C.byref(vo), C.byref(h))
    assert capi.STATUS[st] == "NO_DEVICE" and not h.value


def _panel_from_reference(sc, paths, fasta, bed, tmp_path):
    """A reference panel = the reference program's own STR VCF of these files (EM-trained models), re-indexed with tabix."""
    import gzip
    run_reference(paths, fasta, bed, str(tmp_path / "panel_src.vcf"), 0)
    text = str(tmp_path / "panel.vcf")
    with gzip.open(str(tmp_path / "panel_src.vcf"), "rt") as src, open(text, "w") as dst:
        dst.write(src.read())
    gz = str(tmp_path / "panel.vcf.gz")
    ref = checkers.ref()
    ref.ref_vcf_bgzip_tabix.restype = C.c_int32
    ref.ref_vcf_bgzip_tabix.argtypes = [C.c_char_p, C.c_char_p]
    assert ref.ref_vcf_bgzip_tabix(text.encode(), gz.encode()) == 0
    return gz


@needs_ref
def test_reference_panel_alleles_match_read_vcf_alleles(tmp_path):
    """hipstr_str_vcf_alleles against read_vcf_alleles (src/vcf_input.cpp) on the reference program's own output (CPU only)."""
    from hipstr_b200.capi import StrVcf
    sc = MultiScenario(12, n_regions=4, n_fragments=600)
    paths, fasta, bed = files_of(sc, tmp_path)
    gz = _panel_from_reference(sc, paths, fasta, bed, tmp_path)
    f = checkers.ref().ref_read_vcf_alleles
    f.restype = C.c_int32
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_char_p]
    panel = StrVcf(gz)
    found = 0
    queries = [(s, e, p) for s, e, p in sc.regions] + [(s + 1, e, p) for s, e, p in sc.regions] + [(100, 130, 3), (sc.regions[0][0] - 60, sc.regions[0][0] - 30, 2)]
    for start, stop, period in queries:
        pos, buf = C.c_int32(), C.create_string_buffer(1 << 16)
        rc = f(gz.encode(), b"chr1", start, stop, period, C.byref(pos), len(buf), buf)
        got = panel.alleles("chr1", start, stop)
        if rc == 1:
            assert got == (pos.value, buf.value.decode().splitlines()), (start, stop)
            found += 1
        else:
            assert got is None, (start, stop)
    assert found >= 3
    assert panel.alleles("chr7", 100, 130) is None


@pytest.mark.gpu
@needs_ref
def test_bam_to_vcf_with_reference_panel(tmp_path):
    """--ref-vcf: the alleles of a reference panel are genotyped instead of alleles found in the reads."""
    from hipstr_b200 import capi, pipeline
    sc = MultiScenario(12, n_regions=4, n_fragments=600)
    paths, fasta, bed = files_of(sc, tmp_path)
    gz = _panel_from_reference(sc, paths, fasta, bed, tmp_path)
    s1, e1, p1 = sc.regions[1]
    sc.regions[1] = (s1 + 1, e1, p1)          # the panel has no record with these coordinates: the locus must fail in both
    with open(bed, "w") as fh:
        fh.write(sc.region_text())
    header, want = run_reference(paths, fasta, bed, str(tmp_path / "ref.vcf"), 0, ref_vcf=gz)
    opt = pipeline.Options(min_total_reads=20, ref_vcf=gz)
    with capi.Context(0) as ctx:
        records, summary = pipeline.process_regions(ctx, paths, pipeline.read_fasta(fasta), pipeline.read_regions(bed), opt)
    print(summary)
    assert [canon(r[2]) for r in records] == [canon(w) for w in want] and len(want) == 3
    assert summary["genotype_failed"] == 1


@needs_ref
def test_snp_vcf_edge_cases_match_create_snp_trees(tmp_path):
    """Records the simulator never writes: GT not first in FORMAT, spanning-deletion and symbolic alternates (which htslib counts
    as SNPs), haploid and missing calls, a multi-base REF, no FORMAT column data beyond GT."""
    from hipstr_b200.capi import SnpVcf
    head = ["##fileformat=VCFv4.2", "##contig=<ID=chr1,length=100000>", '##FORMAT=<ID=GT,Number=1,Type=String,Description="g">',
            '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="d">', '##INFO=<ID=AF,Number=A,Type=Float,Description="a">',
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC"]
    rows = [
        (5000, "A", "C", "GT", ["0|1", "1|0", "1|1"]),
        (5010, "G", "T", "DP:GT", ["5:0|1", "7:1|0", "9:0/1"]),
        (5020, "C", "*", "GT", ["0|1", "0|1", "0|0"]),
        (5030, "T", "<X>", "GT:DP", ["0|1:3", "1|0:4", ".|.:0"]),
        (5040, "AC", "A", "GT", ["0|1", "0|1", "0|1"]),
        (5050, "A", "C,G", "GT", ["0|1", "1|2", "0|2"]),
        (5060, "G", "A", "GT", ["0|1", ".", "./."]),
        (5070, "T", "A", "GT", ["1|0", "0|1", "1|0"]),
        (5080, "C", ".", "GT", ["0|0", "0|0", "0|0"]),
        (5090, "a", "g", "GT", ["0|1", "1|0", "0|1"]),
    ]
    text = "\n".join(head + ["chr1\t%d\trs%d\t%s\t%s\t50\tPASS\tAF=0.5\t%s\t%s" % (p, p, r, a, f, "\t".join(g)) for p, r, a, f, g in rows]) + "\n"
    plain, gz = str(tmp_path / "edge.vcf"), str(tmp_path / "edge.vcf.gz")
    with open(plain, "w") as fh:
        fh.write(text)
    ref = checkers.ref()
    ref.ref_vcf_bgzip_tabix.restype = C.c_int32
    ref.ref_vcf_bgzip_tabix.argtypes = [C.c_char_p, C.c_char_p]
    assert ref.ref_vcf_bgzip_tabix(plain.encode(), gz.encode()) == 0
    f = ref.ref_snp_sets
    f.restype = C.c_int32
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    vcf = SnpVcf(gz)
    for start, stop in ((5200, 5230), (4900, 4930), (5044, 5048)):
        buf = C.create_string_buffer(1 << 16)
        n = f(gz.encode(), b"chr1", start, stop, 3, 1000, 15, len(buf), buf)
        assert n >= 0
        off, pos, b1, b2 = vcf.region_sets("chr1", start - 1000, stop + 1000, [(start, stop)], 15)
        lines = []
        for s, name in enumerate(vcf.samples):
            lines.append("S " + name)
            lines += ["%d %s %s" % (pos[k], chr(b1[k]), chr(b2[k])) for k in range(off[s], off[s + 1])]
        assert "\n".join(lines) + "\n" == buf.raw[:n].decode(), (start, stop)
