"""Regenerates tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libhipstr_ref.so (build container only: needs /root/reference via `make -C oracle ref`).

Each fixture = the kwargs of a deterministic synthetic configuration (hipstr_b200.Synth; std::mt19937 seeded,
bit-reproducible across machines) or of a hand-built locus (tests/cases.py) + the reference's outputs for it:
alignment log-likelihoods, posteriors, per-sample normalisers, best diplotypes.  tests/test_golden.py checks the
CPU oracle against them everywhere, and the GPU path against them on the B200 box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases      # noqa: E402
import checkers   # noqa: E402

GOLDEN_SYNTH = ["cfg1_plumbing", "cfg2_shape", "cfg4_shape", "period2", "period1_homopolymer", "period3_noisy", "mates"]
GOLDEN_HAND = [dict(seed=4, flank_opts=(2, 1)), dict(seed=6, homopolymer_edges=True, flank_opts=(2, 2), rep_opts=3, motif="A", copies=9),
               dict(seed=8, motif="AGAT", copies=3, rep_opts=5), dict(seed=10, qual_lo=-5, qual_hi=60)]


def read_ll_from_pools(s, pool_ll):
    out = np.zeros(int(s.read_ll_size))
    off = 0
    for l in range(s.n_loci):
        H = int(s.n_haps[l])
        r0, r1 = s.locus_read_off[l], s.locus_read_off[l + 1]
        pl = pool_ll[s.locus_out_off[l]:s.locus_out_off[l + 1]].reshape(-1, H)
        out[off:off + (r1 - r0) * H] = pl[s.pool_index[r0:r1]].ravel()
        off += (r1 - r0) * H
    return out


def main():
    ref = checkers.ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    for name in GOLDEN_SYNTH:
        s = cases.synth(name)
        ll = checkers.align(ref, "ref_", s.batch, s.n_out)
        read_ll = read_ll_from_pools(s, ll)      # mates are NOT merged here: fixture of the two seams separately
        post, sll, best, tot = checkers.posteriors(ref, "ref_", s.locus_read_off, s.locus_sample_off, s.n_haps, s.haploid,
                                                   read_ll, s.log_p1, s.log_p2, s.sample_label, s.read_weight)
        np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name), kwargs=json.dumps(dict(cases.SYNTH_CASES)[name]),
                            ll=ll, post=post, sample_ll=sll, best=best, total_ll=tot)
        print(name, ll.size, post.size)
    for i, kw in enumerate(GOLDEN_HAND):
        b = cases.handmade_batch(**kw)
        ll = checkers.align(ref, "ref_", b, b.n_out)
        np.savez_compressed(os.path.join(HERE, "hand_%d.npz" % i), kwargs=json.dumps(kw), ll=ll)
        print("hand", i, ll.size)


if __name__ == "__main__":
    main()
# tests/golden/em_params.json (EM stutter learner) is produced the same way: ref_em_train of the compiled reference
# on cases.EM_CASES; see the snippet in the git history of tests/test_em.py / run:
#   python - <<'PY'
#   import json, ctypes as C, cases, checkers
#   from hipstr_b200.capi import EmBatch, c_f64p, c_i32p, c_u8p, em_train
#   f = checkers.ref().ref_em_train; f.restype = C.c_int32
#   f.argtypes = [C.POINTER(EmBatch), C.c_int32, C.c_double, C.c_double, c_f64p, c_u8p, c_i32p, c_f64p]
#   json.dump({n: dict(params=em_train(f, cases.em_case(n)[1])[1].tolist(), converged=em_train(f, cases.em_case(n)[1])[2].tolist())
#              for n, _, _ in cases.EM_CASES}, open("tests/golden/em_params.json", "w"), indent=1)
#   PY
