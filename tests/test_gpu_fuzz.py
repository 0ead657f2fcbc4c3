"""Randomised parity sweep on the GPU: many small hand-built loci with varied motif periods, allele counts, flank
alleles, read lengths (all column-per-lane variants, C = 2..16), seed positions, quality ranges, masks and seedless
reads in ONE batch, against the CPU oracle.  Plus size-independent properties at a BASELINE-like batch size."""
import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import BatchBuilder, Context, Synth

pytestmark = pytest.mark.gpu


def random_locus(rng, long_reads=False):
    motif_len = int(rng.integers(1, 7))
    while True:
        motif = "".join("ACGT"[i] for i in rng.integers(0, 4, motif_len))
        if motif_len == 1 or len(set(motif)) > 1:
            break
    copies = int(rng.integers(2, 18 if not long_reads else 40))
    kw = dict(seed=int(rng.integers(1, 1 << 30)), n_reads=int(rng.integers(3, 14)), motif=motif, copies=copies,
              rep_opts=int(rng.integers(1, 7)), flank_opts=(int(rng.integers(1, 3)), int(rng.integers(1, 3))),
              homopolymer_edges=bool(rng.random() < 0.3), qual_lo=int(rng.choice([-3, 2, 20])), qual_hi=int(rng.choice([41, 41, 55])))
    blocks, reads = cases.handmade(**kw)
    out = []
    for bases, quals, seed in reads:
        if rng.random() < 0.08:
            seed = -1                                   # no usable seed: LL 0 for every haplotype
        elif rng.random() < 0.3:
            cut = int(rng.integers(10, max(11, len(bases) - 5)))   # short read (few columns per lane)
            bases, quals = bases[:cut], quals[:cut]
            seed = int(rng.integers(1, len(bases) - 1))
        out.append((bases, quals, seed))
    return blocks, out


@pytest.mark.parametrize("seed,long_reads", [(1, False), (2, False), (3, True), (4, True)])
def test_fuzz_batch_against_oracle(seed, long_reads):
    rng = np.random.default_rng(seed)
    bb = BatchBuilder()
    n_hap_total, n_pool_total = 0, 0
    for _ in range(40):
        blocks, reads = random_locus(rng, long_reads)
        bb.add_locus(blocks, reads)
        n_hap_total += cases.n_haps_of(blocks)
        n_pool_total += len(reads)
    hap_mask = (rng.random(n_hap_total) < 0.8).astype(np.uint8) if seed % 2 == 0 else None
    pool_mask = (rng.random(n_pool_total) < 0.9).astype(np.uint8) if seed % 2 == 0 else None
    b = bb.build(realign_pool=pool_mask, realign_hap=hap_mask)
    want, wpos = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, want_pos=True, fill=3.5)
    ctx = Context(0)
    got = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 3.5))
    ctx.close()
    d = np.abs(got - want)
    print("[fuzz %d] %d alignments, max|diff| %.3g, not bit-equal %d" % (seed, got.size, d.max(), int((got != want).sum())))
    assert d.max() <= 1e-9


@pytest.mark.parametrize("budget_mb", [1, 3])
def test_fuzz_batch_in_many_chunks(budget_mb, monkeypatch):
    """The same kind of batch with the stutter-table budget forced down to a few megabytes: the tables are computed chunk by
    chunk into two alternating buffers, K1a of a chunk on its own stream while K1b of the previous chunk drains
    (capi.cu run_align).  Results must not depend on the chunking."""
    rng = np.random.default_rng(11)
    bb = BatchBuilder()
    for _ in range(60):
        blocks, reads = random_locus(rng, False)
        bb.add_locus(blocks, reads)
    b = bb.build()
    want, wpos = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, want_pos=True, fill=3.5)
    ctx = Context(0)
    whole = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 3.5))
    n_whole = ctx.lib.hipstr_last_launch_count(ctx.h)
    monkeypatch.setenv("HIPSTR_T_BUDGET_MB", str(budget_mb))
    for rep in range(3):   # repeated calls reuse the two buffers and their events
        got = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 3.5))
        assert np.array_equal(got, whole)
    n_chunked = ctx.lib.hipstr_last_launch_count(ctx.h)
    ctx.close()
    print("[chunks] %d alignments: %d launches in one chunk, %d with a %d MB budget" % (got.size, n_whole, n_chunked, budget_mb))
    assert n_chunked >= 3 * n_whole > 0
    assert np.abs(got - want).max() <= 1e-9


def test_concurrent_contexts_with_different_read_lengths():
    """Four host threads, each with its own context, align batches whose reads differ in length (different shared-memory
    sizes of the same kernels) at the same time, as the pipelines of hipstr_multi_* do: every call succeeds and returns
    what the same batch returns alone.  (Setting a kernel's dynamic shared-memory limit per launch raced here.)"""
    import threading
    batches = []
    for t, long_reads in enumerate([False, True, False, True]):
        rng = np.random.default_rng(100 + t)
        bb = BatchBuilder()
        for _ in range(12):
            blocks, reads = random_locus(rng, long_reads)
            bb.add_locus(blocks, reads)
        batches.append(bb.build())
    alone = []
    for b in batches:
        ctx = Context(0)
        alone.append(ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 3.5)))
        ctx.close()
    errors, results = [], [None] * len(batches)

    def work(i):
        try:
            ctx = Context(0)
            for rep in range(40):
                got = ctx.align_host(batches[i], batches[i].n_out, ll=np.full(batches[i].n_out, 3.5))
                if not np.array_equal(got, alone[i]):
                    raise AssertionError("thread %d, call %d: results differ from the batch run alone" % (i, rep))
            ctx.close()
            results[i] = True
        except Exception as e:   # noqa: BLE001 -- reported below
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert all(results)


def test_properties_at_scale():
    """BASELINE configs[1] shape at 120 loci (1 M alignments; the oracle would need minutes): size-independent checks."""
    s = Synth(n_loci=120, n_samples=100, reads_per_sample=30, n_alleles=8, read_len=150, seed=2000)
    ctx = Context(0)
    S, R = int(s.locus_sample_off[-1]), int(s.n_reads)
    out1 = ctx.genotype_host(s.batch, s.reads_batch(), int(s.read_ll_size), R, int(s.post_size), S, s.n_loci)
    out2 = ctx.genotype_host(s.batch, s.reads_batch(), int(s.read_ll_size), R, int(s.post_size), S, s.n_loci)
    # 1. deterministic: bit-identical from run to run (persistent warps pull jobs in a different order every time)
    for k in ("read_ll", "post", "sample_ll", "best", "total_ll"):
        assert np.array_equal(out1[k], out2[k]), k
    # 2. loci are independent: the first 7 loci alone give exactly the same numbers
    s7 = Synth(n_loci=7, n_samples=100, reads_per_sample=30, n_alleles=8, read_len=150, seed=2000)
    o7 = ctx.genotype_host(s7.batch, s7.reads_batch(), int(s7.read_ll_size), int(s7.n_reads), int(s7.post_size),
                           int(s7.locus_sample_off[-1]), 7)
    assert np.array_equal(o7["read_ll"], out1["read_ll"][:o7["read_ll"].size])
    assert np.array_equal(o7["post"], out1["post"][:o7["post"].size])
    # ... and those 7 loci match the CPU oracle
    want = checkers.align(checkers.oracle(), "oracle_", s7.batch, s7.n_out)
    assert np.array_equal(ctx.align_host(s7.batch, s7.n_out), want)
    # 3. log-likelihoods are finite and negative, identical reads get identical rows
    ll = out1["read_ll"]
    assert np.isfinite(ll).all() and (ll < 0).all()
    H = 8
    rows = ll.reshape(-1, H)
    r0, r1 = s.locus_read_off[0], s.locus_read_off[1]
    pidx = s.pool_index[r0:r1]
    for p in np.unique(pidx)[:50]:
        members = np.nonzero(pidx == p)[0]
        assert (rows[members] == rows[members[0]]).all()
    # 4. posteriors are normalised per sample, the reported best diplotype is their first maximum,
    #    and the locus total is the sum of the sample normalisers
    post = out1["post"].reshape(S, H * H)
    assert np.abs(np.log(np.exp(post).sum(axis=1))).max() < 1e-9
    assert np.array_equal(post.argmax(axis=1), out1["best"][:, 0] * H + out1["best"][:, 1])
    tot = np.add.reduceat(out1["sample_ll"], s.locus_sample_off[:-1])
    assert np.abs(tot - out1["total_ll"]).max() < 1e-6
    # 5. genotype calls recover the simulated truth for most samples
    truth = np.sort(np.ctypeslib.as_array(s.view.true_gt, shape=(S, 2)), axis=1)
    assert (np.sort(out1["best"], axis=1) == truth).all(axis=1).mean() > 0.85
    ctx.close()


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("name,kw", [
    ("configs[2]: 100 samples x 30 reads, 16 alleles", dict(n_loci=3, n_samples=100, reads_per_sample=30, n_alleles=16, read_len=150, seed=3000)),
    ("configs[3]: 500 samples x 5 reads, 32 alleles", dict(n_loci=3, n_samples=500, reads_per_sample=5, n_alleles=32, read_len=150, seed=4000)),
])
def test_full_shapes_against_reference(name, kw):
    """K1 (+ K2 + K3) at the FULL per-locus shapes of BASELINE.json configs[2] and configs[3], three loci each, against the
    compiled, unmodified reference (HapAligner::process_reads + Genotyper::calc_log_sample_posteriors): LLs within the
    north-star tolerance 1e-4 -- and, as everywhere so far, bit-equal -- best diplotypes identical."""
    s = Synth(**kw)
    ref = checkers.ref()
    want = checkers.align(ref, "ref_", s.batch, s.n_out)
    ctx = Context(0)
    got = ctx.align_host(s.batch, s.n_out)
    d = np.abs(got - want)
    print("[%s] %d alignments, max|diff| %.3g, not bit-equal %d" % (name, got.size, d.max(), int((got != want).sum())))
    assert d.max() <= 1e-4 and d.max() <= 1e-9
    # the genotype posteriors on top of them
    reads = s.reads_batch()
    S, R = int(s.locus_sample_off[-1]), int(s.n_reads)
    out = ctx.genotype_host(s.batch, reads, int(s.read_ll_size), R, int(s.post_size), S, s.n_loci)
    ctx.close()
    read_ll = np.concatenate([want[s.locus_out_off[l]:s.locus_out_off[l + 1]].reshape(-1, int(s.n_haps[l]))
                              [s.pool_index[s.locus_read_off[l]:s.locus_read_off[l + 1]]].ravel() for l in range(s.n_loci)])
    post, sll, best, tot = checkers.posteriors(ref, "ref_", s.locus_read_off, s.locus_sample_off, s.n_haps, s.haploid, read_ll, s.log_p1,
                                               s.log_p2, s.sample_label, s.read_weight)
    assert np.array_equal(out["best"].reshape(-1, 2), best.reshape(-1, 2))
    assert np.abs(out["post"] - post).max() <= 1e-9 and np.abs(out["total_ll"] - tot).max() <= 1e-6
