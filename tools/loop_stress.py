"""Stress parity of the whole loop against the unmodified reference SeqStutterGenotyper on many random small loci
(constructor from reads -> genotype(flank assembly on) -> [recompute_stutter_models] -> write_vcf_record).
usage: python tools/loop_stress.py [n_configs] [seed]      (needs a GPU and oracle/_ref/libhipstr_ref.so)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hipstr_b200.capi import Context, Genotyper, Synth
from ref_genotyper import LocusReads, RefGenotyper

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ctx = Context(0)
canon = lambda t: t.replace(":-0.00:", ":0.00:")
n_loci = n_bad = n_failed_loci = n_changed = 0
t0 = time.time()
for c in range(n_cfg):
    period = int(rng.choice([1, 2, 3, 4, 5, 6]))
    kw = dict(n_loci=int(rng.integers(2, 6)), n_samples=int(rng.integers(1, 30)), reads_per_sample=int(rng.integers(1, 25)),
              n_alleles=int(rng.integers(1, 9)), read_len=int(rng.integers(60, 200)), seed=int(rng.integers(1, 1 << 30)), period=period,
              ref_copies=int(rng.integers(max(3, 12 // period), 40 // period + 3)), stutter_rate=float(rng.choice([0.0, 0.05, 0.2, 0.4])),
              sub_rate=float(rng.choice([0.0, 0.005, 0.03])), mate_rate=float(rng.choice([0.0, 0.0, 0.3])),
              flank_snp_freq=float(rng.choice([0.0, 0.0, 0.1, 0.4])), haploid=int(rng.random() < 0.2), trim=int(rng.random() < 0.8))
    recompute = rng.random() < 0.3
    opts = dict(output_gls=int(rng.random() < 0.5), output_pls=int(rng.random() < 0.5), output_phased_gls=int(rng.random() < 0.5),
                output_filters=int(rng.random() < 0.5), output_haplotype_data=int(rng.random() < 0.5))
    min_flank_freq = float(rng.choice([0.01, 0.2]))
    s = Synth(**kw)
    g = Genotyper.from_synth_reads(ctx, s)
    ok = g.genotype(1000, 4, min_flank_freq, True)
    if recompute:
        ok = g.recompute_stutter_models(1000, 4, min_flank_freq)
    reads = [LocusReads(s, l) for l in range(s.n_loci)]
    names = ["S%d" % i for i in range(kw["n_samples"])]
    L = s.n_loci
    loci = g.vcf_loci(["chrS"] * L, ["STR"] * L, [r.region[0] for r in reads], [r.region[1] for r in reads], [period] * L,
                      [r.chrom_seq for r in reads], names * L, names)
    rec = g.write_vcf(loci, **opts)
    for l in range(L):
        n_loci += 1
        r = RefGenotyper(reads[l], reassemble_flanks=True)
        b0 = r.blocks() if r.initialized else None
        want_ok = r.initialized and r.genotype(1000, 4, min_flank_freq)
        if want_ok and recompute:
            want_ok = r.recompute_stutter_models(1000, 4, min_flank_freq)
        problems = []
        if bool(ok[l]) != bool(want_ok):
            problems.append("ok %s vs %s" % (bool(ok[l]), bool(want_ok)))
        elif want_ok:
            if g.blocks(l) != [b[3] for b in r.blocks()]:
                problems.append("blocks")
            else:
                n_changed += [b[3] for b in r.blocks()] != [b[3] for b in b0]
                w, o = r.results(), g.results(l)
                if not np.array_equal(o["best"], w["best"]):
                    problems.append("best")
                if o["read_ll"].size and np.abs(o["read_ll"] - w["read_ll"]).max() > 1e-8:
                    problems.append("ll %.3g" % np.abs(o["read_ll"] - w["read_ll"]).max())
                if canon(rec[l][1]) != canon(r.vcf(**opts).rstrip("\n")):
                    gf, wf = canon(rec[l][1]).split("\t"), canon(r.vcf(**opts).rstrip("\n")).split("\t")
                    diff = [(i, a, b) for i, (a, b) in enumerate(zip(gf, wf)) if a != b]
                    problems.append("vcf %s" % diff[:2])
        else:
            n_failed_loci += 1
        if problems:
            n_bad += 1
            print("MISMATCH cfg %d locus %d %s recompute=%s opts=%s: %s" % (c, l, kw, recompute, opts, problems))
    g.close()
print("stress: %d loci in %d configurations, %d skipped by both, %d with a changed allele set, %d mismatches, %.1f s" %
      (n_loci, n_cfg, n_failed_loci, n_changed, n_bad, time.time() - t0))
