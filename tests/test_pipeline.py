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
                  filters=0):
    f = checkers.ref().ref_process_regions
    f.restype = C.c_int32
    f.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    o = np.array([def_stutter, min_total_reads, remove_dups, require_paired, recalc, gls, pls, filters], np.int32)
    assert f(len(paths), arr, fasta.encode(), bed.encode(), out_vcf.encode(), o.ctypes.data_as(C.POINTER(C.c_int32))) == 0
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
    (4, 0, {}),                                   # stutter models learned by the EM genotyper (K4)
    (5, 1, dict(require_paired=0, gls=1, pls=1, filters=1)),
    (6, 0, dict(recalc=1, remove_dups=0)),
])
def test_bam_to_vcf_matches_reference(seed, def_stutter, kw, tmp_path):
    from hipstr_b200 import capi, pipeline
    sc = MultiScenario(seed, n_regions=4, n_fragments=220 if def_stutter else 600)
    extra = [("chr1", 2000, 2200, 4, 50.0, "TOO_LONG"), ("chr1", 10, 40, 3, 10.0, "CONTIG_END"), ("chr1", 7000, 7030, 3, 10.0, "NO_READS")]
    paths, fasta, bed = files_of(sc, tmp_path, extra)
    header, want = run_reference(paths, fasta, bed, str(tmp_path / "ref.vcf"), def_stutter, **kw)
    opt = pipeline.Options(min_total_reads=20, def_stutter_model=pipeline.DEFAULT_STUTTER if def_stutter else None,
                           recalc_stutter_model=bool(kw.get("recalc", 0)),
                           filter=dict(remove_pcr_dups=kw.get("remove_dups", 1), require_paired_reads=kw.get("require_paired", 1)))
    vcf_opt = dict(output_gls=kw.get("gls", 0), output_pls=kw.get("pls", 0), output_filters=kw.get("filters", 0))
    with capi.Context(0) as ctx:
        records, summary = pipeline.process_regions(ctx, paths, pipeline.read_fasta(fasta), pipeline.read_regions(bed), opt, vcf_opt)
    print(summary)
    assert [canon(r[2]) for r in records] == [canon(w) for w in want]
    assert len(want) >= 2 and summary["too_long"] == 1 and summary["near_contig_end"] == 1 and summary["too_few_reads"] >= 1
    # the sample columns of the header line are the sorted sample names the records were written for
    assert header[-1].split("\t")[9:] == sorted({s for f in sc.files for _, s, _ in f["groups"]})
