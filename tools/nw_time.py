"""Times K6 (batched Needleman-Wunsch, the arithmetic of read left-alignment) against the serial CPU checker.
usage: python tools/nw_time.py [n_pairs]   -- pairs shaped like realign(): ~130-base reads against ~280-base windows."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import checkers
from hipstr_b200.capi import Context
from test_nw import _bind, _call

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(0)
base = []
for k in range(256):
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 280))
    ref = ref[:120] + "AGAT" * 12 + ref[168:]
    a = int(rng.integers(60, 90))
    read = ref[a:a + 130]
    if k % 3 == 0:
        read = read[:45] + "AGAT" + read[45:]
    base.append((ref, read))
ps = [base[i % 256] for i in range(n)]
ctx = Context(0)
refs, reads = [p[0] for p in ps], [p[1] for p in ps]
ctx.nw_align(refs[:1000], reads[:1000])
for rep in range(3):
    t = time.perf_counter()
    ops, score = ctx.nw_align(refs, reads)
    dt = time.perf_counter() - t
    cells = sum(len(a) * len(b) for a, b in ps)
    print("K6: %d alignments in %.1f ms -> %.2f M alignments/s, %.1f G cell updates/s (host strings in, operation strings out)" %
          (n, dt * 1e3, n / dt / 1e6, cells / dt / 1e9))
o = _bind(checkers.oracle(), "oracle_nw_align")
t = time.perf_counter()
m = 2000
for i in range(m):
    _call(o, refs[i], reads[i], False)
dt = time.perf_counter() - t
print("CPU checker (1 core, same algorithm): %.1f k alignments/s" % (m / dt / 1e3))
