"""TEST INFRASTRUCTURE: loaders for the parity checkers under oracle/.

  oracle()  liboracle.so            the CPU restatement (always present after build())
  ref()     _ref/libhipstr_ref.so   the UNMODIFIED reference sources, compiled in the build
                                    container from /root/reference (travels prebuilt to the GPU
                                    box; None when absent)
Nothing under hipstr_b200/ imports this module.
"""
import ctypes as C
import os

import numpy as np

from hipstr_b200.capi import bind_align_abi, c_f64p, c_i32p, c_u8p, ptr, AlignBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libhipstr_ref.so")

_cache = {}


def _load(path, prefix):
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        _cache[path] = None
        return None
    lib = bind_align_abi(C.CDLL(path), prefix)
    f = getattr(lib, prefix + "fast_lse2")
    f.restype = C.c_double
    f.argtypes = [C.c_double, C.c_double]
    f = getattr(lib, prefix + "fast_lse_vec")
    f.restype = C.c_double
    f.argtypes = [c_f64p, C.c_int32]
    _cache[path] = lib
    return lib


def oracle():
    lib = _load(ORACLE_PATH, "oracle_")
    if lib is None:
        raise ImportError("oracle/liboracle.so missing: run `make -C oracle`")
    lib.oracle_hap_options.restype = None
    lib.oracle_hap_options.argtypes = [C.c_int32, c_i32p, C.c_int64, c_i32p]
    lib.oracle_scatter_pool_lls.restype = C.c_int32
    lib.oracle_scatter_pool_lls.argtypes = [C.c_int32, C.c_int32, c_f64p, c_i32p, c_i32p, c_u8p, c_u8p, c_u8p,
                                            c_f64p, c_i32p]
    return lib


def ref():
    lib = _load(REF_PATH, "ref_")
    if lib is not None:
        lib.ref_enumerate_haplotypes.restype = C.c_int32
        lib.ref_enumerate_haplotypes.argtypes = [C.POINTER(AlignBatch), C.c_int32, c_i32p]
    return lib


def align(lib, prefix, batch, n_out, want_pos=False, fill=0.0):
    ll = np.full(n_out, fill, np.float64)
    pos = np.full(n_out, -1, np.int32) if want_pos else None
    st = getattr(lib, prefix + "align_batch")(C.byref(batch), ptr(ll, c_f64p), ptr(pos, c_i32p))
    assert st == 0, st
    return (ll, pos) if want_pos else ll


def posteriors(lib, prefix, locus_read_off, locus_sample_off, n_haps, haploid, read_ll, log_p1, log_p2, sample_label,
               read_weight):
    n_loci = len(n_haps)
    S = int(locus_sample_off[-1])
    post_size = int(sum(int(locus_sample_off[l + 1] - locus_sample_off[l]) * int(n_haps[l]) ** 2 for l in range(n_loci)))
    post = np.zeros(post_size, np.float64)
    sample_ll = np.zeros(S, np.float64)
    best = np.zeros(2 * S, np.int32)
    total = np.zeros(n_loci, np.float64)
    st = getattr(lib, prefix + "posteriors")(
        n_loci, ptr(locus_read_off, c_i32p), ptr(locus_sample_off, c_i32p), ptr(n_haps, c_i32p), ptr(haploid, c_u8p),
        ptr(read_ll, c_f64p), ptr(log_p1, c_f64p), ptr(log_p2, c_f64p), ptr(sample_label, c_i32p),
        ptr(read_weight, c_i32p), ptr(post, c_f64p), ptr(sample_ll, c_f64p), ptr(best, c_i32p), ptr(total, c_f64p))
    assert st == 0, st
    return post, sample_ll, best.reshape(-1, 2), total


def build_hostsim():
    """Builds tests/hostsim/libhipstr_hostsim.so (make decides whether anything is stale) under a file lock: several test
    modules need it and pytest-xdist runs them in different processes."""
    import fcntl
    import subprocess
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "hostsim")
    with open(os.path.join(d, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            subprocess.run(["make", "-s", "-C", d], check=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return os.path.join(d, "libhipstr_hostsim.so")

