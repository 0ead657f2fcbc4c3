"""Golden fixtures generated from the compiled, unmodified reference (tests/golden/make_golden.py).
CPU: the oracle must reproduce them bit for bit (works where /root/reference does not exist).
GPU: the CUDA path must reproduce the log-likelihoods (<= 1e-9, expected bit-equal) and the best diplotypes."""
import glob
import json
import os

import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import Synth

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYNTH = sorted(glob.glob(os.path.join(HERE, "synth_*.npz")))
HAND = sorted(glob.glob(os.path.join(HERE, "hand_*.npz")))


def _read_ll(s, pool_ll):
    out = np.zeros(int(s.read_ll_size))
    off = 0
    for l in range(s.n_loci):
        H = int(s.n_haps[l])
        r0, r1 = s.locus_read_off[l], s.locus_read_off[l + 1]
        pl = pool_ll[s.locus_out_off[l]:s.locus_out_off[l + 1]].reshape(-1, H)
        out[off:off + (r1 - r0) * H] = pl[s.pool_index[r0:r1]].ravel()
        off += (r1 - r0) * H
    return out


def test_fixtures_exist():
    assert len(SYNTH) >= 7 and len(HAND) >= 4


@pytest.mark.parametrize("path", SYNTH, ids=[os.path.basename(p) for p in SYNTH])
def test_oracle_reproduces_reference_fixture(path):
    g = np.load(path)
    s = Synth(**json.loads(str(g["kwargs"])))
    ll = checkers.align(checkers.oracle(), "oracle_", s.batch, s.n_out)
    assert np.array_equal(ll, g["ll"])
    post, sll, best, tot = checkers.posteriors(checkers.oracle(), "oracle_", s.locus_read_off, s.locus_sample_off, s.n_haps,
                                               s.haploid, _read_ll(s, ll), s.log_p1, s.log_p2, s.sample_label, s.read_weight)
    assert np.array_equal(post, g["post"]) and np.array_equal(sll, g["sample_ll"])
    assert np.array_equal(best, g["best"]) and np.array_equal(tot, g["total_ll"])


@pytest.mark.parametrize("path", HAND, ids=[os.path.basename(p) for p in HAND])
def test_oracle_reproduces_handmade_fixture(path):
    g = np.load(path)
    b = cases.handmade_batch(**json.loads(str(g["kwargs"])))
    assert np.array_equal(checkers.align(checkers.oracle(), "oracle_", b, b.n_out), g["ll"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SYNTH + HAND, ids=[os.path.basename(p) for p in SYNTH + HAND])
def test_gpu_reproduces_reference_fixture(path):
    from hipstr_b200.capi import Context
    g = np.load(path)
    kw = json.loads(str(g["kwargs"]))
    ctx = Context(0)
    if "synth_" in path:
        s = Synth(**kw)
        ll = ctx.align_host(s.batch, s.n_out)
        assert np.abs(ll - g["ll"]).max() <= 1e-9
        post, sll, best, tot = ctx.posteriors_host(s.locus_read_off, s.locus_sample_off, s.n_haps, s.haploid, _read_ll(s, ll),
                                                   s.log_p1, s.log_p2, s.sample_label, s.read_weight)
        assert np.abs(post - g["post"]).max() <= 1e-9 and np.array_equal(best, g["best"])
    else:
        b = cases.handmade_batch(**kw)
        assert np.abs(ctx.align_host(b, b.n_out) - g["ll"]).max() <= 1e-9
    ctx.close()
