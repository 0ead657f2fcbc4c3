"""Quick device-time probe of K1 on a resident synthetic batch (not the bench: see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hipstr_b200 as hb

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 50
alleles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = time.time()
s = hb.Synth(n_loci=n_loci, n_samples=100, reads_per_sample=30, n_alleles=alleles, read_len=150, seed=2)
print("synth %.1fs pools=%d out=%d" % (time.time() - t, s.n_pools, s.n_out))
ctx = hb.Context(0)
t = time.time(); h = ctx.upload(s.batch); print("upload %.2fs" % (time.time() - t))
out = torch.zeros(s.n_out, dtype=torch.float64, device="cuda:0")
ctx.enable_timing(True)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for i in range(reps):
    ctx.align_dev(h, out.data_ptr())
    ms = ctx.collect_timing()[0]
    print("run %d: %.2f ms  %.3f M aln/s  launches=%d" % (i, ms, s.n_out / ms / 1e3, ctx.lib.hipstr_last_launch_count(ctx.h)))
t = time.time(); ll = ctx.align_host(s.batch, s.n_out); print("host path %.3fs" % (time.time() - t))
print("checksum", float(out.sum()), float(ll.sum()))
