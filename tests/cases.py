"""Shared test inputs: synthetic configurations and hand-built edge-case batches."""
import numpy as np

from hipstr_b200.capi import BatchBuilder, Synth

# (name, Synth kwargs) -- sized so the CPU oracle finishes each in about a second
SYNTH_CASES = [
    ("cfg1_plumbing", dict(n_loci=1, n_samples=2, reads_per_sample=20, n_alleles=2, read_len=100, seed=1000)),
    ("cfg2_shape", dict(n_loci=2, n_samples=12, reads_per_sample=10, n_alleles=8, read_len=150, seed=2000)),
    ("cfg3_shape", dict(n_loci=2, n_samples=8, reads_per_sample=8, n_alleles=16, read_len=150, seed=3000)),
    ("cfg4_shape", dict(n_loci=2, n_samples=30, reads_per_sample=3, n_alleles=32, read_len=150, seed=4000)),
    ("short_reads", dict(n_loci=3, n_samples=6, reads_per_sample=8, n_alleles=4, read_len=75, seed=5000)),
    ("long_untrimmed", dict(n_loci=2, n_samples=5, reads_per_sample=8, n_alleles=6, read_len=250, seed=5100, trim=0)),
    ("period2", dict(n_loci=3, n_samples=5, reads_per_sample=8, n_alleles=6, read_len=100, seed=5200, period=2, ref_copies=15)),
    ("period1_homopolymer", dict(n_loci=3, n_samples=5, reads_per_sample=8, n_alleles=5, read_len=120, seed=5300, period=1, ref_copies=14)),
    ("period6", dict(n_loci=2, n_samples=5, reads_per_sample=8, n_alleles=5, read_len=200, seed=5400, period=6, ref_copies=6, trim=0)),
    ("period3_noisy", dict(n_loci=2, n_samples=6, reads_per_sample=8, n_alleles=5, read_len=140, seed=5500, period=3, ref_copies=9,
                           stutter_rate=0.3, sub_rate=0.05)),
    ("mates", dict(n_loci=2, n_samples=6, reads_per_sample=6, n_alleles=4, read_len=110, seed=5600, mate_rate=0.5)),
]


def synth(name):
    for n, kw in SYNTH_CASES:
        if n == name:
            return Synth(**kw)
    raise KeyError(name)


def _rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def _rand_qual(rng, n, lo=2, hi=41):
    return "".join(chr(33 + q) for q in rng.integers(lo, hi + 1, n))


def _mutate(rng, s, rate):
    out = []
    for c in s:
        u = rng.random()
        if u < rate / 3:
            continue                          # deletion
        if u < 2 * rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])   # insertion before
        if u > 1 - rate / 3:
            c = "ACGT"[rng.integers(0, 4)]    # substitution
        out.append(c)
    return "".join(out)


def handmade(seed=7, n_reads=24, flank_opts=(1, 1), rep_opts=3, motif="AC", copies=8, n_blocks=3,
             homopolymer_edges=False, qual_lo=2, qual_hi=41, extra_block=False):
    """A locus built directly as blocks + reads, with indels anywhere in the reads and arbitrary seeds.

    flank_opts: number of options of the left / right flank block (flank alleles as after assembly).
    homopolymer_edges: make flanks end/start with runs of the repeat's first/last base so the
    cross-block homopolymer lengths (and the reference's DP-row reuse history) matter.
    extra_block: five blocks flank/repeat/flank/repeat/flank.
    """
    rng = np.random.default_rng(seed)
    lf, rf = _rand_seq(rng, 30), _rand_seq(rng, 30)
    if homopolymer_edges:
        lf = lf[:-3] + motif[0] * 3
        rf = motif[-1] * 2 + rf[2:]
    def variants(s, k):
        outs = [s]
        while len(outs) < k:
            pos = rng.integers(2, len(s) - 2)
            t = s[:pos] + "ACGT"[rng.integers(0, 4)] + s[pos + 1:]
            if t not in outs:
                outs.append(t)
        return outs
    left = variants(lf, flank_opts[0])
    right = variants(rf, flank_opts[1])
    if homopolymer_edges and flank_opts[0] > 1:
        left[1] = left[1][:-1] + ("G" if motif[0] != "G" else "T")   # one left allele does NOT end in the run
    reps = [motif * (copies + d) for d in range(rep_opts)]
    if homopolymer_edges:
        reps = [motif[0] * (1 + i) + r for i, r in enumerate(reps)]   # alleles start with runs of different length
    blocks = [(0, left), (len(motif), reps), (0, right)]
    if extra_block:
        mid = _rand_seq(rng, 12)
        blocks = [(0, left), (len(motif), reps), (0, [mid]), (3, ["GAT" * 5, "GAT" * 6]), (0, right)]
    reads = []
    for _ in range(n_reads):
        hap = "".join(opts[rng.integers(0, len(opts))] for _, opts in blocks)
        s = rng.integers(0, 12)
        e = len(hap) - rng.integers(0, 12)
        rd = _mutate(rng, hap[s:e], 0.03)
        seedpos = int(rng.integers(1, len(rd) - 1))
        reads.append((rd, _rand_qual(rng, len(rd), qual_lo, qual_hi), seedpos))
    return blocks, reads


def handmade_batch(realign_pool=None, realign_hap=None, **kw):
    bb = BatchBuilder()
    blocks, reads = handmade(**kw)
    bb.add_locus(blocks, reads)
    return bb.build(realign_pool=realign_pool, realign_hap=realign_hap)


def n_haps_of(blocks):
    h = 1
    for _, opts in blocks:
        h *= len(opts)
    return h


# ---- EM stutter learner inputs (BASELINE.json configs[3] shape, scaled down) --------------------------
EM_CASES = [
    ("em_diploid", dict(n_loci=3, n_samples=40, reads_per_sample=8, n_alleles=6, read_len=150, seed=4100, stutter_rate=0.15), 0),
    ("em_many_alleles", dict(n_loci=2, n_samples=60, reads_per_sample=5, n_alleles=12, read_len=150, seed=4200, stutter_rate=0.3), 0),
    ("em_haploid_p2", dict(n_loci=2, n_samples=30, reads_per_sample=6, n_alleles=5, read_len=150, seed=4300, stutter_rate=0.2,
                           period=2, ref_copies=15), 1),
    ("em_cfg4_shape", dict(n_loci=2, n_samples=120, reads_per_sample=5, n_alleles=32, read_len=150, seed=4400, stutter_rate=0.1), 0),
]


def em_case(name, out_of_frame=0.03):
    """Returns (Synth, hipstr_em_batch_t).  Read STR sizes come from the simulated reads (what ExtractCigar would
    report); a few reads get a +/-1 bp out-of-frame artefact so every bucket of the M-step is exercised."""
    from hipstr_b200.capi import make_em_batch
    for n, kw, haploid in EM_CASES:
        if n == name:
            s = Synth(**kw)
            period = kw.get("period", 4)
            ref_bp = period * kw.get("ref_copies", 12)
            rng = np.random.default_rng(kw["seed"])
            diff = np.ctypeslib.as_array(s.view.read_bp_diff, shape=(s.n_reads,)).copy()
            jitter = rng.random(s.n_reads) < out_of_frame
            diff[jitter] += rng.choice([-1, 1], int(jitter.sum()))
            p1 = np.log(rng.uniform(0.2, 1.0, s.n_reads))
            p2 = np.log(rng.uniform(0.2, 1.0, s.n_reads))
            b = make_em_batch(s.locus_read_off, s.locus_sample_off, diff + ref_bp, s.sample_label, p1, p2,
                              np.full(s.n_loci, period), np.full(s.n_loci, ref_bp), np.full(s.n_loci, haploid))
            return s, b
    raise KeyError(name)
