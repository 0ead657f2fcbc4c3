"""Times the C++ multi-GPU driver (hipstr_multi_*) on synthetic loci: loci/s of create_from_reads -> genotype -> write_vcf.
usage: multi_time.py LOCI ALLELES PIPELINES WINDOW [DEVICES]"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hipstr_b200 as hb
from hipstr_b200.capi import Genotyper, MultiGenotyper

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 400
alleles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pipelines = int(sys.argv[3]) if len(sys.argv) > 3 else 2
window = int(sys.argv[4]) if len(sys.argv) > 4 else 100
devices = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
t = time.time()
s = hb.Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=alleles, read_len=150, seed=2000)
print("synth %.1fs" % (time.time() - t))
names = ["S%d" % i for i in range(100)]
cl = int(s.view.chrom_len)
raw = C.string_at(s.view.chrom_seqs, n_loci * cl)
loci = Genotyper.vcf_loci(["chr1"] * n_loci, ["STR%d" % l for l in range(n_loci)], [s.view.region_start] * n_loci,
                          [s.view.region_stop] * n_loci, [4] * n_loci, [raw[l * cl:(l + 1) * cl] for l in range(n_loci)],
                          names * n_loci, names)
m = MultiGenotyper(devices=devices, pipelines=pipelines)
for rep in range(3):
    t = time.perf_counter()
    ok, rec = m.genotype_synth(s, loci, window)
    dt = time.perf_counter() - t
    st = m.stats()
    print("rep %d: %d loci in %.2fs -> %.1f loci/s; ok %d records %d; %s" % (rep, n_loci, dt, n_loci / dt, int(ok.sum()),
                                                                            sum(r is not None for r in rec), st))
m.close()
