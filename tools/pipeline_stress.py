"""Stress parity of BAM files -> VCF records against the unmodified reference program on many random inputs and option sets.
usage: python tools/pipeline_stress.py [n_seeds] [first_seed]      (needs a GPU and oracle/_ref/libhipstr_ref.so)"""
import os
import pathlib
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_pipeline as T
from hipstr_b200 import capi, pipeline
from ingest_sim import MultiScenario

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
canon = lambda t: t.replace(":-0.00:", ":0.00:")
ctx = capi.Context(0)
n_regions = n_records = n_bad = 0
t0 = time.time()
for seed in range(first, first + n_seeds):
    rng = np.random.default_rng(seed)
    def_stutter = int(rng.random() < 0.5)
    kw = dict(require_paired=int(rng.random() < 0.6), remove_dups=int(rng.random() < 0.7), recalc=int(rng.random() < 0.3),
              gls=int(rng.random() < 0.5), pls=int(rng.random() < 0.5), filters=int(rng.random() < 0.5), haploid=int(rng.random() < 0.2))
    use_snps = rng.random() < 0.6
    sc = MultiScenario(seed, n_regions=int(rng.integers(2, 6)), n_files=int(rng.integers(1, 4)),
                       n_fragments=int(rng.integers(150, 400)) if def_stutter else int(rng.integers(450, 800)))
    with tempfile.TemporaryDirectory() as tmp:
        tp = pathlib.Path(tmp)
        paths, fasta, bed = T.files_of(sc, tp)
        snp_vcf = T.write_snp_vcf(sc, tp, seed)[1] if use_snps else None
        header, want = T.run_reference(paths, fasta, bed, str(tp / "ref.vcf"), def_stutter, snp_vcf=snp_vcf, **kw)
        opt = pipeline.Options(min_total_reads=20, snp_vcf=snp_vcf, def_stutter_model=pipeline.DEFAULT_STUTTER if def_stutter else None,
                               recalc_stutter_model=bool(kw["recalc"]), haploid_chroms=("chr1",) if kw["haploid"] else (),
                               filter=dict(remove_pcr_dups=kw["remove_dups"], require_paired_reads=kw["require_paired"]))
        vcf_opt = dict(output_gls=kw["gls"], output_pls=kw["pls"], output_filters=kw["filters"])
        records, summary = pipeline.process_regions(ctx, paths, pipeline.read_fasta(fasta), pipeline.read_regions(bed), opt, vcf_opt)
    got = [canon(r[2]) for r in records]
    want = [canon(w) for w in want]
    n_regions += len(sc.regions)
    n_records += len(want)
    if got != want:
        n_bad += 1
        print("MISMATCH seed", seed, def_stutter, kw, use_snps, len(got), len(want))
        for a, b in zip(got, want):
            if a != b:
                fa, fb = a.split("\t"), b.split("\t")
                print("  first differing column:", next((i for i, (x, y) in enumerate(zip(fa, fb)) if x != y), None), fa[:2])
                break
print("pipeline stress: %d seeds, %d regions, %d records, %d mismatching runs, %.1f s" % (n_seeds, n_regions, n_records, n_bad, time.time() - t0))
