"""K6: batched Needleman-Wunsch, the arithmetic of read left-alignment.
CPU: the oracle restatement equals the compiled reference's NeedlemanWunsch::Align (operation strings and scores) on
random pairs with indels, repeats, N bases, free and penalised reference ends.
GPU: hipstr_nw_align_batch_host must return exactly the oracle's strings and scores."""
import ctypes as C

import numpy as np
import pytest

import checkers

needs_ref = pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libhipstr_ref.so not built")


def _bind(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_int32
    f.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(C.c_float)]
    return f


def _call(f, ref, read, penalty):
    buf = C.create_string_buffer(len(ref) + len(read) + 2)
    score = C.c_float()
    n = f(ref.encode(), len(ref), read.encode(), len(read), int(penalty), buf, C.byref(score))
    return (buf.value.decode() if n >= 0 else None), score.value


def pairs(seed, n, ref_lo=40, ref_hi=300, with_n=True):
    """(window, read) pairs the way realign() sees them: the read is a mutated slice of the window, which often holds a
    short tandem repeat so that equally good gap placements exist and the tie rules matter."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        L = int(rng.integers(ref_lo, ref_hi))
        ref = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
        if k % 2 == 0:   # plant a repeat
            motif = "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(1, 7))))
            at = int(rng.integers(5, max(6, L - 30)))
            rep = motif * int(rng.integers(3, 12))
            ref = (ref[:at] + rep + ref[at:])[:max(L, at + len(rep) + 5)]
        a = int(rng.integers(0, max(1, len(ref) // 3)))
        b = len(ref) - int(rng.integers(0, max(1, len(ref) // 3)))
        read = list(ref[a:b])
        for _ in range(int(rng.integers(0, 6))):   # substitutions, insertions, deletions (often whole motif copies)
            if not read:
                break
            p = int(rng.integers(0, len(read)))
            kind = rng.integers(0, 3)
            if kind == 0:
                read[p] = "ACGT"[rng.integers(0, 4)]
            elif kind == 1:
                read[p:p] = list("ACGT"[rng.integers(0, 4)] * int(rng.integers(1, 9)))
            else:
                del read[p:p + int(rng.integers(1, 9))]
        if with_n and k % 7 == 0 and read:
            read[int(rng.integers(0, len(read)))] = "N"
        if len(read) < 2:
            read = list("AC")
        out.append((ref, "".join(read)))
    out += [("ACGTACGTAC", "ACGTACGTAC"), ("AAAAAAAAAA", "AAA"), ("ACACACACACACAC", "ACACACAC"), ("GATTACA", "TTTTTTTTT"),
            ("A", "A"), ("AC", "G"), ("ACGT", "ACGTACGT")]
    return out


@needs_ref
@pytest.mark.parametrize("penalty", [False, True])
def test_oracle_nw_equals_reference(penalty):
    o, r = _bind(checkers.oracle(), "oracle_nw_align"), _bind(checkers.ref(), "ref_nw_align")
    n_gapped = 0
    for ref, read in pairs(1, 400):
        got, want = _call(o, ref, read, penalty), _call(r, ref, read, penalty)
        assert got == want, (ref, read)
        n_gapped += ("I" in want[0]) or ("D" in want[0].strip("D"))
    assert n_gapped > 100


@pytest.mark.gpu
@pytest.mark.parametrize("penalty", [False, True])
def test_gpu_nw_equals_oracle(penalty):
    from hipstr_b200.capi import Context
    o = _bind(checkers.oracle(), "oracle_nw_align")
    ps = pairs(2, 1500) + pairs(3, 40, ref_lo=300, ref_hi=420)
    ctx = Context(0)
    ops, score = ctx.nw_align([p[0] for p in ps], [p[1] for p in ps], penalty)
    ctx.close()
    bad = 0
    for (ref, read), got, sc in zip(ps, ops, score):
        want, wsc = _call(o, ref, read, penalty)
        if got != want or sc != np.float32(wsc):
            if bad == 0:
                print("first mismatch\n ref  %s\n read %s\n gpu  %s %r\n want %s %r" % (ref, read, got, sc, want, wsc))
            bad += 1
    assert bad == 0, "%d of %d alignments differ" % (bad, len(ps))
