"""Parity of the CUDA path (through the C-ABI of include/hipstr_b200.h) against the CPU oracle, on a
real B200.  Bar: bit-exact for integer outputs (seed placement, best diplotype); log-likelihoods
within 1e-4 as BASELINE.json's north_star states -- the flank DP and the approximate log-sum-exp
replicas are in fact designed to be bit-identical, so the tests also REPORT the number of values
that are not bit-equal and assert a much tighter 1e-9."""
import ctypes as C

import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import BatchBuilder, Context, load

pytestmark = pytest.mark.gpu

LL_TOL = 1e-4      # north_star tolerance on log-likelihood fields
TIGHT = 1e-9       # what the design is expected to deliver


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def _report(name, got, want):
    d = np.abs(got - want)
    bad = int((got != want).sum())
    print("[%s] n=%d max|diff|=%.3g not-bit-equal=%d" % (name, got.size, d.max() if d.size else 0.0, bad))
    if d.size and d.max() > TIGHT:
        i = int(np.argmax(d))
        print("   worst index %d: gpu=%r oracle=%r" % (i, got[i], want[i]))
        idx = np.nonzero(d > TIGHT)[0][:10]
        print("   first bad indices", idx, got[idx], want[idx])
    return d.max() if d.size else 0.0


@pytest.mark.parametrize("name", [n for n, _ in cases.SYNTH_CASES])
def test_align_synthetic(ctx, name):
    s = cases.synth(name)
    want, wpos = checkers.align(checkers.oracle(), "oracle_", s.batch, s.n_out, want_pos=True)
    got, gpos = ctx.align_host(s.batch, s.n_out, want_pos=True)
    worst = _report(name, got, want)
    assert worst <= LL_TOL
    assert worst <= TIGHT
    assert np.array_equal(gpos, wpos)
    assert ctx.lib.hipstr_last_launch_count(ctx.h) >= 1


@pytest.mark.parametrize("kw", [
    dict(seed=1), dict(seed=2, motif="ACG", copies=6), dict(seed=3, motif="A", copies=12, rep_opts=4),
    dict(seed=4, flank_opts=(2, 1)), dict(seed=5, flank_opts=(2, 3), rep_opts=2),
    dict(seed=6, homopolymer_edges=True, flank_opts=(2, 2), rep_opts=3, motif="A", copies=9),
    dict(seed=7, homopolymer_edges=True, rep_opts=4, motif="AT", copies=7),
    dict(seed=8, motif="AGAT", copies=3, rep_opts=5), dict(seed=9, motif="ACGTAC", copies=1, rep_opts=2),
    dict(seed=10, qual_lo=-5, qual_hi=60), dict(seed=11, extra_block=True),
    dict(seed=12, n_reads=40, motif="AAAG", copies=20, rep_opts=3),
], ids=lambda k: "-".join("%s=%s" % i for i in k.items()))
def test_align_handmade(ctx, kw):
    b = cases.handmade_batch(**kw)
    want, wpos = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, want_pos=True)
    got, gpos = ctx.align_host(b, b.n_out, want_pos=True)
    worst = _report(str(kw), got, want)
    assert worst <= TIGHT
    assert np.array_equal(gpos, wpos)


@pytest.mark.parametrize("kw", [dict(seed=21, flank_opts=(2, 2), rep_opts=3, homopolymer_edges=True, motif="A", copies=9),
                                dict(seed=22, rep_opts=6), dict(seed=23, flank_opts=(3, 1), rep_opts=2)],
                         ids=["homop", "rep6", "flank3"])
def test_align_masks_leave_entries_untouched(ctx, kw):
    blocks, reads = cases.handmade(**kw)
    H = cases.n_haps_of(blocks)
    rng = np.random.default_rng(kw["seed"])
    hap_mask = (rng.random(H) < 0.6).astype(np.uint8)
    hap_mask[rng.integers(0, H)] = 1
    pool_mask = (rng.random(len(reads)) < 0.7).astype(np.uint8)
    b = BatchBuilder().add_locus(blocks, reads).build(realign_pool=pool_mask, realign_hap=hap_mask)
    want = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, fill=123.25)
    got = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 123.25))
    assert _report("masks", got, want) <= TIGHT
    g2 = got.reshape(len(reads), H)
    assert np.all(g2[pool_mask == 0] == 123.25) and np.all(g2[:, hap_mask == 0] == 123.25)


def test_seedless_pool_and_multi_locus(ctx):
    bb = BatchBuilder()
    per_locus = []
    for sd in (31, 32, 33):
        blocks, reads = cases.handmade(seed=sd, rep_opts=2 + sd % 3, n_reads=10 + sd % 5)
        reads[2] = (reads[2][0], reads[2][1], -1)
        bb.add_locus(blocks, reads)
        per_locus.append((cases.n_haps_of(blocks), len(reads)))
    b = bb.build()
    want = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, fill=5.0)
    got = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 5.0))
    assert _report("multi-locus", got, want) <= TIGHT
    off = 0
    for H, P in per_locus:
        assert np.all(got[off + 2 * H: off + 3 * H] == 0.0)
        off += H * P


def test_empty_batch_and_bad_args(ctx):
    b = BatchBuilder().build()
    got = ctx.align_host(b, 0)
    assert got.size == 0
    blocks, reads = cases.handmade(seed=41)
    reads[0] = (reads[0][0], reads[0][1], 0)     # seed on the first base: the reference dies
    bad = BatchBuilder().add_locus(blocks, reads).build()
    with pytest.raises(Exception) as e:
        ctx.align_host(bad, bad.n_out)
    assert "INVALID_SEED" in str(e.value)
    rep_first = BatchBuilder().add_locus([(2, ["ACACAC"]), (0, ["ACGTACGT"])], [("ACACGT", "IIIIII", 2)]).build()
    with pytest.raises(Exception) as e:
        ctx.align_host(rep_first, rep_first.n_out)
    assert "UNSUPPORTED" in str(e.value)


def test_resident_batch_matches_host_path(ctx):
    import torch
    s = cases.synth("cfg2_shape")
    want = ctx.align_host(s.batch, s.n_out)
    h = ctx.upload(s.batch)
    out = torch.zeros(s.n_out, dtype=torch.float64, device="cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.align_dev(h, out.data_ptr())
    torch.cuda.synchronize()
    ctx.set_stream(0)
    assert np.array_equal(out.cpu().numpy(), want)
    ctx.free_batch(h)
    assert load().hipstr_batch_num_alignments(C.byref(s.batch)) == int(((s.pool_seed >= 0).astype(np.int64) * np.repeat(s.n_haps, np.diff(s.locus_pool_off))).sum())


def test_scatter_and_mate_merge(ctx):
    o = checkers.oracle()
    from hipstr_b200.capi import c_f64p, c_i32p, c_u8p, ptr
    rng = np.random.default_rng(3)
    for H, R, P in [(1, 7, 3), (8, 300, 120), (27, 64, 64)]:
        pool_ll = -rng.uniform(1, 300, P * H)
        pool_seed = rng.integers(-1, 90, P).astype(np.int32)
        pool_index = rng.integers(0, P, R).astype(np.int32)
        second = (rng.random(R) < 0.3).astype(np.uint8)
        second[0] = 0
        for copy_read, hap_mask in [(None, None), ((rng.random(R) < 0.7).astype(np.uint8), (rng.random(H) < 0.7).astype(np.uint8))]:
            base = -rng.uniform(1, 300, R * H)
            want, wseed = base.copy(), np.full(R, -7, np.int32)
            o.oracle_scatter_pool_lls(R, H, ptr(pool_ll, c_f64p), ptr(pool_seed, c_i32p), ptr(pool_index, c_i32p),
                                      ptr(second, c_u8p), ptr(copy_read, c_u8p), ptr(hap_mask, c_u8p), ptr(want, c_f64p),
                                      ptr(wseed, c_i32p))
            got, gseed = base.copy(), np.full(R, -7, np.int32)
            ctx.scatter_host(H, pool_ll, pool_seed, pool_index, second, got, gseed, copy_read, hap_mask)
            assert np.array_equal(got, want)
            assert np.array_equal(gseed, wseed)


@pytest.mark.parametrize("name", ["cfg1_plumbing", "cfg2_shape", "cfg4_shape", "mates"])
def test_posteriors(ctx, name):
    s = cases.synth(name)
    pool_ll = checkers.align(checkers.oracle(), "oracle_", s.batch, s.n_out)
    # pool -> read on the host for the test input (K2 is covered above)
    read_ll = np.zeros(int(s.read_ll_size))
    off = 0
    for l in range(s.n_loci):
        H = int(s.n_haps[l])
        r0, r1 = s.locus_read_off[l], s.locus_read_off[l + 1]
        pl = pool_ll[s.locus_out_off[l]:s.locus_out_off[l + 1]].reshape(-1, H)
        read_ll[off:off + (r1 - r0) * H] = pl[s.pool_index[r0:r1]].ravel()
        off += (r1 - r0) * H
    rng = np.random.default_rng(11)
    for haploid, phased in [(0, False), (0, True), (1, False)]:
        p1 = np.log(rng.uniform(0.05, 1.0, s.n_reads)) if phased else s.log_p1
        p2 = np.log(rng.uniform(0.05, 1.0, s.n_reads)) if phased else s.log_p2
        hap = np.full(s.n_loci, haploid, np.uint8)
        args = (s.locus_read_off, s.locus_sample_off, s.n_haps, hap, read_ll, p1, p2, s.sample_label, s.read_weight)
        wp, ws, wb, wt = checkers.posteriors(checkers.oracle(), "oracle_", *args)
        gp, gs, gb, gt = ctx.posteriors_host(*args)
        finite = wp > -1e300
        assert np.array_equal(finite, gp > -1e300)
        print("[post %s haploid=%d phased=%d] max|dpost|=%.3g max|dsll|=%.3g" %
              (name, haploid, phased, np.abs(gp[finite] - wp[finite]).max(), np.abs(gs - ws).max()))
        assert np.abs(gp[finite] - wp[finite]).max() <= 1e-9      # exp/log of CUDA vs glibc differ in ulps
        assert np.abs(gs - ws).max() <= 1e-9 and np.abs(gt - wt).max() <= 1e-8
        assert np.array_equal(gb, wb)


def _oracle_genotype(s, copy_read=None, read_ll0=None):
    """align -> scatter -> posteriors with the CPU oracle, locus by locus."""
    from hipstr_b200.capi import c_f64p, c_i32p, c_u8p, ptr
    o = checkers.oracle()
    fill = 0.0
    pool_ll = checkers.align(o, "oracle_", s.batch, s.n_out, fill=fill)
    read_ll = np.zeros(int(s.read_ll_size)) if read_ll0 is None else read_ll0.copy()
    read_seed = np.full(s.n_reads, -2, np.int32)
    off = 0
    hap_mask_all = s.batch.realign_hap
    for l in range(s.n_loci):
        H = int(s.n_haps[l])
        r0, r1 = int(s.locus_read_off[l]), int(s.locus_read_off[l + 1])
        p0, p1 = int(s.locus_pool_off[l]), int(s.locus_pool_off[l + 1])
        pl = np.ascontiguousarray(pool_ll[s.locus_out_off[l]:s.locus_out_off[l + 1]])
        seeds = np.ascontiguousarray(s.pool_seed[p0:p1])
        rl = np.ascontiguousarray(read_ll[off:off + (r1 - r0) * H])
        rs = np.ascontiguousarray(read_seed[r0:r1])
        cr = None if copy_read is None else np.ascontiguousarray(copy_read[r0:r1])
        hm = None
        if hap_mask_all:
            hm = np.ascontiguousarray(np.ctypeslib.as_array(hap_mask_all, shape=(int(s.locus_hap_off[-1]),))[s.locus_hap_off[l]:s.locus_hap_off[l + 1]])
        o.oracle_scatter_pool_lls(r1 - r0, H, ptr(pl, c_f64p), ptr(seeds, c_i32p),
                                  ptr(np.ascontiguousarray(s.pool_index[r0:r1]), c_i32p),
                                  ptr(np.ascontiguousarray(s.second_mate[r0:r1]), c_u8p), ptr(cr, c_u8p), ptr(hm, c_u8p),
                                  ptr(rl, c_f64p), ptr(rs, c_i32p))
        read_ll[off:off + (r1 - r0) * H] = rl
        read_seed[r0:r1] = rs
        off += (r1 - r0) * H
    post, sll, best, tot = checkers.posteriors(o, "oracle_", s.locus_read_off, s.locus_sample_off, s.n_haps, s.haploid,
                                               read_ll, s.log_p1, s.log_p2, s.sample_label, s.read_weight)
    return dict(read_ll=read_ll, read_seed=read_seed, post=post, sample_ll=sll, best=best, total_ll=tot)


@pytest.mark.parametrize("name", ["cfg1_plumbing", "cfg2_shape", "mates", "period3_noisy"])
def test_genotype_batch_host(ctx, name):
    s = cases.synth(name)
    want = _oracle_genotype(s)
    got = ctx.genotype_host(s.batch, s.reads_batch(), int(s.read_ll_size), int(s.n_reads), int(s.post_size),
                            int(s.locus_sample_off[-1]), s.n_loci)
    assert np.array_equal(got["read_ll"], want["read_ll"])
    assert np.array_equal(got["read_seed"], want["read_seed"])
    assert np.abs(got["post"] - want["post"]).max() <= 1e-9
    assert np.abs(got["sample_ll"] - want["sample_ll"]).max() <= 1e-9
    assert np.abs(got["total_ll"] - want["total_ll"]).max() <= 1e-8
    assert np.array_equal(got["best"], want["best"])
    h2d, d2h, launches = ctx.traffic()
    assert h2d > 0 and d2h >= got["read_ll"].nbytes and launches >= 3


def test_genotype_batch_resident_and_copy_read_mask(ctx):
    import torch
    s = cases.synth("mates")
    rng = np.random.default_rng(1)
    # resident path == host path
    want = ctx.genotype_host(s.batch, s.reads_batch(), int(s.read_ll_size), int(s.n_reads), int(s.post_size),
                             int(s.locus_sample_off[-1]), s.n_loci)
    h = ctx.upload_genotype(s.batch, s.reads_batch())
    S = int(s.locus_sample_off[-1])
    t = dict(read_ll=torch.zeros(int(s.read_ll_size), dtype=torch.float64, device="cuda:0"),
             read_seed=torch.zeros(int(s.n_reads), dtype=torch.int32, device="cuda:0"),
             post=torch.zeros(int(s.post_size), dtype=torch.float64, device="cuda:0"),
             sample_ll=torch.zeros(S, dtype=torch.float64, device="cuda:0"),
             best=torch.zeros(2 * S, dtype=torch.int32, device="cuda:0"),
             total_ll=torch.zeros(s.n_loci, dtype=torch.float64, device="cuda:0"))
    ctx.genotype_dev(h, *[t[k].data_ptr() for k in ("read_ll", "read_seed", "post", "sample_ll", "best", "total_ll")])
    torch.cuda.synchronize()
    ctx.free_genotype(h)
    for k in ("read_ll", "read_seed", "post", "sample_ll", "total_ll"):
        assert np.array_equal(t[k].cpu().numpy(), want[k]), k
    assert np.array_equal(t["best"].cpu().numpy().reshape(-1, 2), want["best"])
    # copy_read mask: excluded reads keep the caller's read_ll / read_seed content
    copy_read = (rng.random(s.n_reads) < 0.7).astype(np.uint8)
    base = -rng.uniform(1, 50, int(s.read_ll_size))
    ref = _oracle_genotype(s, copy_read=copy_read, read_ll0=base)
    got = ctx.genotype_host(s.batch, s.reads_batch(copy_read), int(s.read_ll_size), int(s.n_reads), int(s.post_size),
                            S, s.n_loci, read_ll=base.copy(), read_seed=np.full(s.n_reads, -2, np.int32))
    assert np.array_equal(got["read_ll"], ref["read_ll"])
    assert np.array_equal(got["read_seed"], ref["read_seed"])
    assert np.array_equal(got["best"], ref["best"])


def test_seedless_long_read_in_short_read_batch(ctx):
    """Seedless pools ride in the shortest-read launch; their bases must not be staged into its (smaller) landing
    zone (regression: out-of-bounds TMA write found by the configs[4] sweep under compute-sanitizer)."""
    bb = BatchBuilder()
    blocks, reads = cases.handmade(seed=51, n_reads=6)
    short = [(r[0][:40], r[1][:40], 20) for r in reads[:3]]
    long_seedless = [(reads[3][0] * 3, reads[3][1] * 3, -1)]
    bb.add_locus(blocks, short + long_seedless + [(reads[4][0], reads[4][1], reads[4][2])])
    b = bb.build()
    want = checkers.align(checkers.oracle(), "oracle_", b, b.n_out, fill=7.0)
    got = ctx.align_host(b, b.n_out, ll=np.full(b.n_out, 7.0))
    assert _report("seedless-long", got, want) <= TIGHT
