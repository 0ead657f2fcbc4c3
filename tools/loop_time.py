"""Times the whole seam-B1 pipeline on synthetic loci: constructor from reads -> genotype() -> write_vcf_record.
usage: python tools/loop_time.py [n_loci] [n_alleles] [reassemble 0/1] [flank_snp_freq]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hipstr_b200
from hipstr_b200.capi import Context, Genotyper, Synth

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_alleles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
assemble = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
snp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
t = time.time()
s = Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=n_alleles, read_len=150, seed=2000, flank_snp_freq=snp)
print("synth %.2fs: %d reads" % (time.time() - t, s.n_reads))
ctx = Context(0)
for rep in range(2):
    t0 = time.time()
    g = Genotyper.from_synth_reads(ctx, s)
    t1 = time.time()
    ok = g.genotype(1000, 4, 0.01, assemble)
    t2 = time.time()
    names = ["S%d" % i for i in range(100)]
    raw = __import__("ctypes").string_at(s.view.chrom_seqs, n_loci * s.view.chrom_len)
    cl = s.view.chrom_len
    loci = g.vcf_loci(["chr1"] * n_loci, ["STR%d" % l for l in range(n_loci)], [s.view.region_start] * n_loci, [s.view.region_stop] * n_loci,
                      [4] * n_loci, [raw[l * cl:(l + 1) * cl] for l in range(n_loci)], names * n_loci, names)
    rec = g.write_vcf(loci)
    t3 = time.time()
    st = g.stats()
    print("rep %d: construct %.2fs genotype %.2fs vcf %.2fs total %.2fs -> %.1f loci/s; ok %d/%d; %s" %
          (rep, t1 - t0, t2 - t1, t3 - t2, t3 - t0, n_loci / (t3 - t0), int(ok.sum()), n_loci, st))
    print("   stages:", {k: round(v, 3) for k, v in g.timing().items()})
    print("   decide by phase:", {k: round(v, 3) for k, v in g.phase_timing().items()})
    g.close()
