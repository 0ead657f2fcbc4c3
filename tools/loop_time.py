"""Times the whole seam-B1 pipeline on synthetic loci: constructor from reads -> genotype() -> write_vcf_record.
usage: python tools/loop_time.py [n_loci] [n_alleles] [reassemble 0/1] [flank_snp_freq] [pipelines]
pipelines > 1 splits the loci into that many windows, each driven by its own thread + context, so that the host
stages of one window overlap the device stages of another."""
import ctypes
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hipstr_b200
from hipstr_b200.capi import Context, Genotyper, Synth

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_alleles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
assemble = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
snp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
pipes = int(sys.argv[5]) if len(sys.argv) > 5 else 1
t = time.time()
s = Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=n_alleles, read_len=150, seed=2000, flank_snp_freq=snp)
print("synth %.2fs: %d reads" % (time.time() - t, s.n_reads))
ctxs = [Context(0) for _ in range(pipes)]
names = ["S%d" % i for i in range(100)]
cl = s.view.chrom_len
raw = ctypes.string_at(s.view.chrom_seqs, n_loci * cl)


def run(k, out):
    l0, l1 = k * n_loci // pipes, (k + 1) * n_loci // pipes
    L = l1 - l0
    t0 = time.time()
    g = Genotyper.from_synth_reads(ctxs[k], s, loci_range=(l0, l1))
    t1 = time.time()
    ok = g.genotype(1000, 4, 0.01, assemble)
    t2 = time.time()
    loci = g.vcf_loci(["chr1"] * L, ["STR%d" % l for l in range(l0, l1)], [s.view.region_start] * L, [s.view.region_stop] * L,
                      [4] * L, [raw[l * cl:(l + 1) * cl] for l in range(l0, l1)], names * L, names)
    rec = g.write_vcf(loci)
    t3 = time.time()
    out[k] = dict(construct=t1 - t0, genotype=t2 - t1, vcf=t3 - t2, ok=int(ok.sum()), stats=g.stats(), stages=g.timing(),
                  phases=g.phase_timing(), records=sum(r is not None for r in rec))
    g.close()


for rep in range(3):
    out = [None] * pipes
    t0 = time.time()
    threads = [threading.Thread(target=run, args=(k, out)) for k in range(pipes)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    dt = time.time() - t0
    print("rep %d: %d pipeline(s), total %.2fs -> %.1f loci/s; ok %d/%d records %d" %
          (rep, pipes, dt, n_loci / dt, sum(o["ok"] for o in out), n_loci, sum(o["records"] for o in out)))
    for k, o in enumerate(out):
        print("   [%d] construct %.2f genotype %.2f vcf %.2f %s" % (k, o["construct"], o["genotype"], o["vcf"], o["stats"]))
        print("       stages:", {a: round(b, 3) for a, b in o["stages"].items()})
    if pipes == 1:
        print("       trace call parts (cumulative over reps):", {a: round(b, 3) for a, b in ctxs[0].trace_seconds().items()})
        print("       decide by phase:", {a: round(b, 3) for a, b in out[0]["phases"].items()})
