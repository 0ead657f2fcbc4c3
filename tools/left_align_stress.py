"""Stress parity of the batched left alignment (K6 + host loop) against the reference's TrimAlignment / convertAlignment /
realign + reuse-by-sequence loop on many random loci.  usage: python tools/left_align_stress.py [n_configs] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import checkers
from hipstr_b200.capi import Context, LeftAligned, Synth, make_locus_reads
from test_left_align import _bind_ref, python_loop, raw_reads, ref_one, trim_like_reference

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ref = _bind_ref(checkers.ref())
ctx = Context(0)
n_reads = n_bad = n_nw = n_dropped = 0
t0 = time.time()
for c in range(n_cfg):
    period = int(rng.choice([1, 2, 3, 4, 5, 6]))
    kw = dict(n_loci=int(rng.integers(1, 4)), n_samples=int(rng.integers(1, 12)), reads_per_sample=int(rng.integers(1, 15)),
              n_alleles=int(rng.integers(1, 9)), read_len=int(rng.integers(60, 220)), seed=int(rng.integers(1, 1 << 30)), period=period,
              ref_copies=int(rng.integers(max(3, 12 // period), 40 // period + 3)), stutter_rate=float(rng.choice([0.0, 0.1, 0.4])),
              sub_rate=float(rng.choice([0.0, 0.01, 0.05])), trim=int(rng.random() < 0.7))
    s = Synth(**kw)
    reads, chroms, lro, t_lo, t_hi = raw_reads(s, seed=c, clip_rate=float(rng.choice([0.0, 0.2, 0.5])), lower_rate=0.2)
    trim = rng.random() < 0.7
    R = len(reads)
    raw = make_locus_reads(lro, s.locus_sample_off, reads, s.sample_label, np.arange(R), s.log_p1, s.log_p2, s.haploid)
    la = LeftAligned(ctx, s.n_loci, raw, chroms, [t_lo] * s.n_loci if trim else None, [t_hi] * s.n_loci if trim else None)
    got, got_lro = la.reads()
    want = []
    for l in range(s.n_loci):
        locus_raw = reads[lro[l]:lro[l + 1]]
        per_read = [ref_one(ref, rd, chroms[l], (t_lo, t_hi) if trim else None) for rd in locus_raw]
        trimmed = [trim_like_reference(rd, t_lo, t_hi) if trim else (rd[2], rd[3]) for rd in locus_raw]
        for i, aln in python_loop(per_read, trimmed):
            want.append((lro[l] + i, aln))
    n_reads += R
    n_nw += la.nw_alignments
    n_dropped += R - len(want)
    if [int(x) for x in la.source] != [w[0] for w in want] or got != [w[1] for w in want]:
        n_bad += 1
        print("MISMATCH cfg %d %s trim=%s" % (c, kw, trim))
    la.close()
print("left-align stress: %d reads in %d configurations, %d Needleman-Wunsch alignments on the GPU, %d reads dropped by both, "
      "%d configurations with a mismatch, %.1f s" % (n_reads, n_cfg, n_nw, n_dropped, n_bad, time.time() - t0))
