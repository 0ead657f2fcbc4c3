"""Times K5 (traceback) end to end on a synthetic batch: one trace per pooled read against a random haplotype."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hipstr_b200 as hb
from test_trace import block_starts
n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 20
s = hb.Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=8, read_len=150, seed=2000)
rng = np.random.default_rng(0)
pools = np.nonzero(s.pool_seed >= 0)[0].astype(np.int32)
loc = np.searchsorted(s.locus_pool_off, pools, side="right") - 1
haps = rng.integers(0, 8, len(pools)).astype(np.int32)
ctx = hb.Context(0)
bs = block_starts(s.batch)
ctx.trace(s.batch, bs, pools[:64], haps[:64])
for rep in range(2):
    before = np.array(list(ctx.trace_seconds().values()))
    t = time.perf_counter()
    out = ctx.trace(s.batch, bs, pools, haps)
    dt = time.perf_counter() - t
    inside = np.array(list(ctx.trace_seconds().values())) - before
    print("    inside hipstr_trace_batch_host: lowering %.1f ms, ordering + uploads %.1f ms, kernels %.1f ms, downloads %.1f ms -> %.2f M traces/s "
          "(the rest of the wall time is this script's ctypes / numpy wrapper)" % (*(1e3 * inside), len(pools) / max(inside.sum(), 1e-9) / 1e6))
    print("K5: %d traces in %.1f ms -> %.2f M traces/s (host buffers, end to end); stutter!=0 in %d" %
          (len(pools), dt * 1e3, len(pools) / dt / 1e6, int(((out["stutter_size"][:, 1] != 0)).sum())))
# small-batch latency: what one locus of the loop asks for between two rounds (a batch of ONE locus)
s1 = hb.Synth(n_loci=1, n_samples=100, reads_per_sample=30, n_alleles=8, read_len=150, seed=2001)
pools1 = np.nonzero(s1.pool_seed >= 0)[0].astype(np.int32)
haps1 = rng.integers(0, 8, len(pools1)).astype(np.int32)
bs1 = block_starts(s1.batch)
for n_small in (16, 64, 256):
    ctx.trace(s1.batch, bs1, pools1[:n_small], haps1[:n_small])
    t = time.perf_counter()
    for _ in range(50):
        ctx.trace(s1.batch, bs1, pools1[:n_small], haps1[:n_small])
    print("K5 small batch: %d traces of one locus, %.3f ms per call (host buffers in, host buffers out, includes the ctypes wrapper)" %
          (n_small, (time.perf_counter() - t) / 50 * 1e3))
