"""EM stutter learner (SURVEY.md 8 a17 / seam B4).
CPU: oracle restatement == EMStutterGenotyper::train of the compiled reference, bit for bit, and == the committed
golden parameters.  GPU: K4 (through hipstr_em_train_host) vs the oracle: the learner mixes exact exp/log (CUDA libm
vs glibc differ in the last ulps) with the bit-faithful approximate log-sum-exps, so the bar is 1e-6 absolute on the
six model parameters (they are printed with 6 significant digits by StutterModel::write), 1e-9 relative on the final
log-likelihood, and identical convergence flag / iteration count."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import cases
import checkers
from hipstr_b200.capi import EmBatch, c_f64p, c_i32p, c_u8p, em_train

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "em_params.json")
NAMES = [n for n, _, _ in cases.EM_CASES]


def _fn(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_int32
    f.argtypes = [C.POINTER(EmBatch), C.c_int32, C.c_double, C.c_double, c_f64p, c_u8p, c_i32p, c_f64p]
    return f


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")
@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_reference_em(name):
    _, b = cases.em_case(name)
    st, po, co, io, lo = em_train(_fn(checkers.oracle(), "oracle_em_train"), b)
    assert st == 0
    st, pr, cr, _, _ = em_train(_fn(checkers.ref(), "ref_em_train"), b)
    assert st == 0
    assert np.array_equal(po, pr) and np.array_equal(co, cr)
    assert np.all(po > 0) and np.all(po[:, [0, 3]] <= 0.999)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_em(name):
    gold = json.load(open(GOLDEN))[name]
    _, b = cases.em_case(name)
    st, po, co, io, lo = em_train(_fn(checkers.oracle(), "oracle_em_train"), b)
    assert np.array_equal(po, np.array(gold["params"])) and list(co) == gold["converged"]


def test_oracle_em_iteration_cap():
    _, b = cases.em_case("em_diploid")
    st, po, co, io, lo = em_train(_fn(checkers.oracle(), "oracle_em_train"), b, max_iter=2)
    assert np.all(io == 2) and not co.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_em_matches_oracle(name):
    from hipstr_b200.capi import Context
    _, b = cases.em_case(name)
    st, po, co, io, lo = em_train(_fn(checkers.oracle(), "oracle_em_train"), b)
    ctx = Context(0)
    pg, cg, ig, lg = ctx.em_train(b)
    print("[em %s] max|dparam|=%.3g iters gpu=%s oracle=%s max rel dLL=%.3g" %
          (name, np.abs(pg - po).max(), ig, io, np.abs((lg - lo) / lo).max()))
    assert np.abs(pg - po).max() <= 1e-6
    assert np.array_equal(cg, co) and np.array_equal(ig, io)
    assert np.abs((lg - lo) / lo).max() <= 1e-9
    gold = json.load(open(GOLDEN))[name]
    assert np.abs(pg - np.array(gold["params"])).max() <= 1e-6
    # iteration cap: train() returns false and keeps the last M-step's model
    _, c2, i2, _ = ctx.em_train(b, max_iter=2)
    assert np.all(i2 == 2) and not c2.any()
    ctx.close()
