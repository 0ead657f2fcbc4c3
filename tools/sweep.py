"""BASELINE.json configs[4]: allele-count sweep 2-64 x read-length sweep 75-250 bp, 100 samples x 30 reads, on one
B200.  Reads are NOT trimmed to STR +/- 40 bp (with a 48-bp STR the trim makes read_len > ~130 a no-op, SURVEY.md
8d), so the read length really grows.  Also times the EM stutter learner (K4) on the configs[3] shape.
Prints a markdown table; the run under gpurun is committed as profiles/r1_sweep.md."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import hipstr_b200 as hb

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = hb.Context(0)
ctx.enable_timing(True)
print("| alleles | read len | pools/locus | alignments | K1 ms | M aln/s | G cell-updates/s |")
print("|---:|---:|---:|---:|---:|---:|---:|")
for alleles in (2, 4, 8, 16, 32, 64):
    for read_len in (75, 100, 150, 200, 250):
        s = hb.Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=alleles, read_len=read_len, seed=5000 + alleles,
                     trim=0)
        h = ctx.upload(s.batch)
        out = torch.zeros(s.n_out, dtype=torch.float64, device="cuda:0")
        for _ in range(2):
            ctx.align_dev(h, out.data_ptr())
        ctx.collect_timing()
        for _ in range(3):
            ctx.align_dev(h, out.data_ptr())
        ms = ctx.collect_timing()[0] / 3
        ctx.free_batch(h)
        lens = np.diff(s.pool_seq_off).astype(np.int64)
        hap_len = 35 + 58 + 35
        cells = float(((lens - 1) * hap_len * np.repeat(s.n_haps, np.diff(s.locus_pool_off))).sum())
        print("| %d | %d | %d | %d | %.2f | %.2f | %.1f |" % (int(s.n_haps[0]), read_len, s.n_pools // n_loci, s.n_out, ms, s.n_out / ms / 1e3,
                                                          cells / ms / 1e6))
        s.close()

# EM stutter learner, configs[3] shape: 500 samples x 5 reads, 32 requested alleles
import cases
from hipstr_b200.capi import make_em_batch
for n_em in (50, 200):
    s = hb.Synth(n_loci=n_em, n_samples=500, reads_per_sample=5, n_alleles=32, read_len=150, seed=4000, stutter_rate=0.1)
    diff = np.ctypeslib.as_array(s.view.read_bp_diff, shape=(s.n_reads,))
    b = make_em_batch(s.locus_read_off, s.locus_sample_off, diff + 48, s.sample_label, s.log_p1, s.log_p2, np.full(n_em, 4),
                      np.full(n_em, 48), np.zeros(n_em))
    ctx.em_train(b)
    t = time.perf_counter()
    prm, conv, it, ll = ctx.em_train(b)
    dt = time.perf_counter() - t
    print("\nEM (K4): %d loci x 2500 reads, %d alleles: %.1f ms end to end (host buffers), %.0f loci/s, iterations mean %.1f, converged %d/%d"
          % (n_em, len(set(diff[:2500].tolist()) | {0}), dt * 1e3, n_em / dt, it.mean(), int(conv.sum()), n_em))
