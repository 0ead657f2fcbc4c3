"""The HOST side of the hot path without a GPU: the lockstep genotyping loop (constructor from reads, genotype() with
allele discovery / pruning / flank assembly, write_vcf_record) and the multi-GPU window dealer, run over
tests/hostsim/libhipstr_hostsim.so -- the product's host sources with the device entry points simulated on the CPU
oracle (test infrastructure, see tests/hostsim/hostsim.cpp).  Compared with the UNMODIFIED reference SeqStutterGenotyper
where oracle/_ref is built.  The GPU versions of the same comparisons are tests/test_genotyper.py and
tests/test_vcf_record.py."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

import checkers
from hipstr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libhipstr_hostsim.so")
needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")


@pytest.fixture(scope="module")
def sim():
    """capi bound to the host simulation for the duration of this module."""
    checkers.build_hostsim()
    saved = (capi._lib, capi.LIB_PATH)
    capi._lib, capi.LIB_PATH = None, HOSTSIM
    try:
        yield capi.load()
    finally:
        capi._lib, capi.LIB_PATH = saved


def synth(n_loci, seed, **kw):
    import hipstr_b200 as hb
    args = dict(n_samples=6, reads_per_sample=14, n_alleles=3, read_len=100, stutter_rate=0.25, flank_snp_freq=0.3)
    args.update(kw)
    return hb.Synth(n_loci=n_loci, seed=seed, **args)


def vcf_loci(s):
    S = int(s.locus_sample_off[1])
    names = ["S%d" % i for i in range(S)]
    cl = int(s.view.chrom_len)
    raw = C.string_at(s.view.chrom_seqs, s.n_loci * cl)
    return capi.Genotyper.vcf_loci(["chrS"] * s.n_loci, ["STR"] * s.n_loci, [s.view.region_start] * s.n_loci,
                                   [s.view.region_stop] * s.n_loci, [int(s.cfg.period) or 4] * s.n_loci,
                                   [raw[l * cl:(l + 1) * cl] for l in range(s.n_loci)], names * s.n_loci, names)


def single_context_records(s):
    ctx = capi.Context(0)
    g = capi.Genotyper.from_synth_reads(ctx, s)
    ok = g.genotype(1000, 4, 0.01, True)
    rec = g.write_vcf(vcf_loci(s))
    g.close()
    ctx.close()
    return ok, rec


def norm(text):
    return text.rstrip("\n").replace(":-0.00:", ":0.00:")


@needs_ref
@pytest.mark.parametrize("seed", [3, 11])
def test_host_loop_matches_reference(sim, seed):
    """Same allele sets, genotypes and VCF text as the reference's SeqStutterGenotyper, locus by locus."""
    from ref_genotyper import LocusReads, RefGenotyper
    s = synth(3, seed)
    ok, rec = single_context_records(s)
    changed = 0
    for l in range(s.n_loci):
        r = RefGenotyper(LocusReads(s, l), reassemble_flanks=True)
        assert r.genotype() == bool(ok[l])
        if ok[l]:
            assert norm(rec[l][1]) == norm(r.vcf())
            changed += 1
    assert changed > 0


@pytest.mark.parametrize("workers,window", [((0,), 1), ((0, 0), 2), ((0, 0, 0), 5)])
def test_multi_dealer_equals_single_context(sim, workers, window):
    """Whatever the number of workers and the window size, the per-locus records equal the single-context ones."""
    s = synth(7, 21)
    ok1, rec1 = single_context_records(s)
    m = capi.MultiGenotyper(devices=workers, pipelines=2)
    ok2, rec2 = m.genotype_synth(s, vcf_loci(s), window)
    st = m.stats()
    m.close()
    assert ok1.tolist() == ok2.tolist() and rec1 == rec2
    assert sum(st["windows_per_worker"]) == -(-s.n_loci // window) and st["alignments"] > 0


def test_multi_window_order_is_heaviest_first(sim):
    s = synth(9, 5, reads_per_sample=10)
    m = capi.MultiGenotyper(devices=(0,), pipelines=1)
    order = m.window_order(s, 2)
    m.close()
    cost = [int(s.locus_read_off[min(s.n_loci, 2 * (w + 1))] - s.locus_read_off[2 * w]) for w in range(5)]
    assert sorted(order.tolist()) == list(range(5))
    assert all(cost[order[i]] >= cost[order[i + 1]] for i in range(4))


def test_multi_shared_dealer_across_handles(sim):
    """Two handles (two processes under torchrun) pull windows from ONE counter: every window is processed exactly once and
    the union of their records is the single-context result."""
    s = synth(8, 33)
    ok1, rec1 = single_context_records(s)
    lock, state = threading.Lock(), {"next": 0}

    def dealer():
        with lock:
            k = state["next"]
            state["next"] += 1
            return k
    handles = [capi.MultiGenotyper(devices=(0,), pipelines=1) for _ in range(2)]
    results = [None, None]

    def run(i):
        results[i] = handles[i].genotype_synth(s, vcf_loci(s), 3, next_window=dealer)
    threads = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    merged = [a if a is not None else b for a, b in zip(results[0][1], results[1][1])]
    assert all(not (a is not None and b is not None) for a, b in zip(results[0][1], results[1][1]))
    assert merged == rec1
    assert ((results[0][0] | results[1][0]) == ok1).all()
    for h in handles:
        h.close()


@pytest.mark.timeout(300)
@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")   # forking with parked threads is the point
def test_host_thread_pool_under_concurrent_and_nested_use():
    """hipstr::parallel_run (one process-wide pool of parked threads behind every parallel loop of the host side): eight callers
    at once, 200 rounds each, nested calls from inside every index; every index runs exactly once and nothing deadlocks.  Then
    the same from a forked child (a fork keeps none of the parked threads: the child starts a pool of its own)."""
    lib = C.CDLL(HOSTSIM)
    f = lib.hostsim_exercise_thread_pool
    f.restype = C.c_int64
    f.argtypes = [C.c_int32] * 4
    assert f(8, 200, 37, 5) == 0
    assert f(1, 50, 1, 4) == 0          # fewer indices than workers
    assert f(3, 20, 1000, 16) == 0
    pid = os.fork()
    if pid == 0:
        try:
            ok = f(4, 50, 37, 5) == 0
        except BaseException:
            ok = False
        os._exit(0 if ok else 1)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0

