#!/usr/bin/env python
"""bench.py -- read x haplotype HMM alignments/s of the HipSTR hot path on B200.

One "step" = one pass of the hot path (K1 alignment of every pooled read against every candidate
haplotype, K2 pool->read scatter + mate merge, K3 genotype posteriors) over one batch of synthetic
loci of BASELINE.json configs[1]: 1 000 loci, 100 samples x 30 reads, 8 candidate alleles, 150-bp
reads (SURVEY.md 8d generator).  Under torchrun every rank owns its own 1 000 loci (weak scaling:
loci are independent, no data-path collective) and the per-locus genotype records are gathered to
rank 0 over NCCL at the end of every step.

  value   alignments/s with the batch resident in HBM (hipstr_genotype_batch_dev)
  e2e     the same through the host-buffer C-ABI call (hipstr_genotype_batch_host): host flattening,
          H2D of the inputs, K1-K3, D2H of read LLs / posteriors / genotypes inside the timed region
  roofline, cpu_baseline, clocks: see DESIGN.md "Measurement"
  full_loop  (N=1) loci/s through the whole seam-B1 path on the same loci -- constructor from reads, genotype() with
          allele discovery / pruning / flank assembly rounds, write_vcf_record -- next to cpu_baseline.full_loop, the
          unmodified reference SeqStutterGenotyper on one locus per host core

--impl reference times the reference's own CPU code (oracle/_ref/libhipstr_ref.so compiled from the
unmodified sources; the C++ restatement in oracle/ if that library is absent) on the host cores.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hipstr_b200 as hb                      # noqa: E402
from hipstr_b200.capi import c_f64p, c_i32p, c_u8p, ptr  # noqa: E402

METRIC = "read x haplotype HMM alignments/sec"
UNIT = "alignments/s"


# ---------------------------------------------------------------------------------------------
# CPU side: the reference (or the oracle port) over host cores.  Test/bench infrastructure only.
# ---------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_lib():
    import checkers
    lib = checkers.ref()
    if lib is not None:
        return lib, "ref_", "reference"
    return checkers.oracle(), "oracle_", "port"


def _cpu_worker(rng):
    """Align + posteriors for loci [l0, l1) of the inherited synthetic batch; returns seconds."""
    l0, l1 = rng
    s, lib, prefix = _CPU["synth"], _CPU["lib"], _CPU["prefix"]
    t0 = time.perf_counter()
    ll = np.zeros(int(s.locus_out_off[l1] - s.locus_out_off[l0]), np.float64)
    base = int(s.locus_out_off[l0])
    # the align entry points index ll_out by locus_out_off, so hand them a pointer rebased to locus 0
    ll_ptr = C.cast(ll.ctypes.data - 8 * base, c_f64p)
    st = getattr(lib, prefix + "align_loci")(C.byref(s.batch), l0, l1, ll_ptr, None)
    assert st == 0
    # pool -> read (seq_stutter_genotyper.cpp:530-548); no mates in the bench workload
    parts = []
    for l in range(l0, l1):
        H = int(s.n_haps[l])
        pl = ll[int(s.locus_out_off[l]) - base:int(s.locus_out_off[l + 1]) - base].reshape(-1, H)
        parts.append(pl[s.pool_index[s.locus_read_off[l]:s.locus_read_off[l + 1]]].ravel())
    read_ll = np.concatenate(parts)
    S = int(s.locus_sample_off[l1] - s.locus_sample_off[l0])
    post = np.zeros(int(sum(int(s.locus_sample_off[l + 1] - s.locus_sample_off[l]) * int(s.n_haps[l]) ** 2
                            for l in range(l0, l1))), np.float64)
    sll = np.zeros(int(s.locus_sample_off[-1]), np.float64)
    best = np.zeros(2 * int(s.locus_sample_off[-1]), np.int32)
    tot = np.zeros(l1 - l0, np.float64)
    st = getattr(lib, prefix + "posteriors")(
        l1 - l0, ptr(s.locus_read_off[l0:], c_i32p), ptr(s.locus_sample_off[l0:], c_i32p), ptr(s.n_haps[l0:], c_i32p),
        ptr(s.haploid[l0:], c_u8p), ptr(read_ll, c_f64p), ptr(s.log_p1, c_f64p), ptr(s.log_p2, c_f64p),
        ptr(s.sample_label, c_i32p), ptr(s.read_weight, c_i32p), ptr(post, c_f64p), ptr(sll, c_f64p),
        ptr(best, c_i32p), ptr(tot, c_f64p))
    assert st == 0 and S >= 0
    return time.perf_counter() - t0


def _cpu_loop_worker(l):
    """The reference's whole per-locus path (constructor with haplotype generation, genotype() with flank assembly,
    write_vcf_record) for locus l of the inherited synthetic batch; returns (seconds, genotype() succeeded)."""
    from ref_genotyper import LocusReads, RefGenotyper
    rd = LocusReads(_CPU["synth"], l)
    t0 = time.perf_counter()
    g = RefGenotyper(rd, reassemble_flanks=True)
    ok = g.initialized and g.genotype(1000, 4, 0.01)
    if ok:
        g.vcf()
    dt = time.perf_counter() - t0
    g.close()
    return dt, bool(ok)


def cpu_full_loop(synth, cores):
    """One locus per core through the unmodified reference SeqStutterGenotyper (bounded sample)."""
    import checkers
    if checkers.ref() is None:
        return None
    _CPU.update(synth=synth)
    n = min(synth.n_loci, cores)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(n) as pool:
        res = pool.map(_cpu_loop_worker, range(n))
    wall = time.perf_counter() - t0
    return {"loci_per_s": n / wall, "cores": n, "seconds_per_locus_per_core": float(np.mean([r[0] for r in res])),
            "sample": "first %d loci, one per core: constructor + genotype(flank assembly on) + write_vcf_record, %.1f s wall" % (n, wall)}


def gpu_full_loop(device, synth, pipelines, gather=None, locus_base=0):
    """Seam B1 end to end on the GPU: hipstr_genotyper_create_from_reads -> genotype (flank assembly on) -> write_vcf.
    The loci are split into `pipelines` windows, each driven by its own host thread and context, so that the host stages
    of one window (per-locus decisions, trace stitching, VCF text) overlap the device stages of another."""
    from hipstr_b200.capi import Context, Genotyper
    L = synth.n_loci
    names = ["S%d" % i for i in range(int(synth.locus_sample_off[1]))]
    cl = synth.view.chrom_len
    raw = C.string_at(synth.view.chrom_seqs, L * cl)
    period = int(synth.cfg.period) or 4
    ctxs = [Context(device) for _ in range(pipelines)]

    # the inputs of write_vcf_record (region descriptors, chromosome pointers, sample names) are host buffers the caller
    # owns, like the reads: built once, outside the timed region
    bounds = [(k * L // pipelines, (k + 1) * L // pipelines) for k in range(pipelines)]
    vcf_inputs = [Genotyper.vcf_loci(["chr1"] * (l1 - l0), ["STR%d" % l for l in range(l0, l1)], [synth.view.region_start] * (l1 - l0),
                                     [synth.view.region_stop] * (l1 - l0), [period] * (l1 - l0),
                                     [raw[l * cl:(l + 1) * cl] for l in range(l0, l1)], names * (l1 - l0), names) for l0, l1 in bounds]

    def window(k, out):
        l0, l1 = bounds[k]
        g = Genotyper.from_synth_reads(ctxs[k], synth, loci_range=(l0, l1))
        ok = g.genotype(1000, 4, 0.01, True)
        rec = g.write_vcf(vcf_inputs[k])
        out[k] = (int(ok.sum()), sum(r is not None for r in rec), g.stats(), g.timing(),
                  [(locus_base + l0 + i, "chr1", r[0], r[1]) for i, r in enumerate(rec) if r is not None])
        g.close()

    best = None
    for rep in range(3):   # the first pass warms the allocations
        out = [None] * pipelines
        t0 = time.perf_counter()
        threads = [threading.Thread(target=window, args=(k, out)) for k in range(pipelines)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if gather is not None:   # N > 1: the one collective, finished records to rank 0 (inside the timed region)
            dt = gather([r for o in out for r in o[4]], t0)
        else:
            dt = time.perf_counter() - t0
        stages = {}
        for o in out:
            for k, v in o[3].items():
                stages[k] = round(stages.get(k, 0.0) + v, 4)
        aln = sum(o[2]["alignments"] for o in out)
        res = {"loci_per_s": L / dt, "seconds": dt, "loci": L, "pipelines": pipelines, "loci_genotyped": sum(o[0] for o in out),
               "records": sum(o[1] for o in out), "alignments": aln, "traces": sum(o[2]["traces"] for o in out),
               "rounds": max(o[2]["rounds"] for o in out), "alignments_per_s": aln / dt,
               "stage_seconds_summed_over_windows": stages,
               "host_threads": int(os.environ.get("HIPSTR_HOST_THREADS", host_cores())),
               "what": "hipstr_genotyper_create_from_reads + genotype(1000, 4, 0.01, reassemble_flanks) + write_vcf, host buffers in, VCF text out"}
        if best is None or res["loci_per_s"] > best["loci_per_s"]:
            best = res
    for c in ctxs:
        c.close()
    return best


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_pass(synth, cores, loci_per_core):
    """One bounded CPU sample: `cores` forked workers, `loci_per_core` loci each.  Returns
    (alignments, wall seconds, sum of worker seconds)."""
    lib, prefix, _ = _cpu_lib()
    _CPU.update(synth=synth, lib=lib, prefix=prefix)
    n = min(synth.n_loci, cores * loci_per_core)
    cores = min(cores, n)
    bounds = [(i * n // cores, (i + 1) * n // cores) for i in range(cores)]
    aln = int(((synth.pool_seed[:synth.locus_pool_off[n]] >= 0).astype(np.int64) *
               np.repeat(synth.n_haps[:n], np.diff(synth.locus_pool_off[:n + 1]))).sum())
    t0 = time.perf_counter()
    if cores == 1:
        secs = [_cpu_worker(bounds[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            secs = pool.map(_cpu_worker, bounds)
    return aln, time.perf_counter() - t0, float(sum(secs)), n, cores


# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(s):
    """bytes K1 must move per pass = reads (1-byte bases + quals + 16 B record) +
    haplotypes (2 B per base + 64 B record) + repeat-allele tables + one 8-byte LL per alignment."""
    b = s.batch
    L = b.n_loci
    opt_off = np.ctypeslib.as_array(b.opt_seq_off, shape=(b.n_options + 1,))
    blk_off = np.ctypeslib.as_array(b.block_opt_off, shape=(b.n_blocks + 1,))
    period = np.ctypeslib.as_array(b.block_period, shape=(b.n_blocks,))
    lbo = np.ctypeslib.as_array(b.locus_block_off, shape=(L + 1,))
    opt_len = np.diff(opt_off).astype(np.int64)
    read_bytes = int((2 * np.diff(s.pool_seq_off).astype(np.int64) + 16).sum())
    hap_bytes = rep_bytes = 0
    for l in range(L):
        H, hlen = int(s.n_haps[l]), 0
        for k in range(lbo[l], lbo[l + 1]):
            lens = opt_len[blk_off[k]:blk_off[k + 1]]
            hlen += int(lens.mean())
            if period[k] > 0:
                ndel = np.minimum(6, lens // period[k])
                rep_bytes += int((13 * 8 + (ndel + 1) * 4 * lens).sum())
        hap_bytes += H * (2 * hlen + 64)
    n_aln = hb.load().hipstr_batch_num_alignments(C.byref(b))
    return read_bytes + hap_bytes + rep_bytes + 8 * n_aln, n_aln


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        self.on = False

    def _pump(self):
        for line in self.proc.stdout:
            if self.on:
                self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, warnings) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--reads-per-sample", type=int, default=30)
    ap.add_argument("--alleles", type=int, default=8)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-loop", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3 if a.impl == "ours" else 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # the ranks of one box share its host cores: split them instead of oversubscribing (read by the library)
        os.environ.setdefault("HIPSTR_HOST_THREADS", str(max(2, host_cores() // world)))
    workload = "%d synthetic loci, %d samples x %d reads, %d alleles, %d bp reads (BASELINE.json configs[1])" % (
        a.loci, a.samples, a.reads_per_sample, a.alleles, a.read_len)
    config = {"workload": workload, "loci_per_gpu": a.loci, "sharding": "independent loci per rank, NCCL gather of per-locus genotype records per step",
              "l2": "inputs+outputs of a step (~0.6 GB) exceed the 126 MB L2; no explicit flush"}

    if a.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        s = hb.Synth(n_loci=min(a.loci, 2 * cores), n_samples=a.samples, reads_per_sample=a.reads_per_sample,
                     n_alleles=a.alleles, read_len=a.read_len, seed=2000)
        _, _, kind = _cpu_lib()
        for _ in range(a.warmup):
            cpu_pass(s, cores, 1)
        tot_aln, tot_t = 0, 0.0
        for _ in range(a.steps):
            aln, wall, _, n, used = cpu_pass(s, cores, 1)
            tot_aln += aln
            tot_t += wall
        v = tot_aln / tot_t
        sample = "%d loci per step (1 per core) of the same synthetic workload" % n
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    t_gen = time.time()
    s = hb.Synth(n_loci=a.loci, n_samples=a.samples, reads_per_sample=a.reads_per_sample, n_alleles=a.alleles,
                 read_len=a.read_len, seed=2000 + rank)
    t_gen = time.time() - t_gen
    alg_bytes, n_aln = algorithmic_bytes(s)

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:   # before CUDA is initialised: workers are forked
        cores = host_cores()
        _, _, kind = _cpu_lib()
        per_core = 2 if cores <= 16 else 1
        aln, wall, cpu_s, n, used = cpu_pass(s, cores, per_core)
        cpu_baseline = {"value": aln / wall, "unit": UNIT, "cores": used, "kind": kind,
                        "sample": "first %d loci of the workload, align + posteriors, %d forked workers, %.1f s wall / %.1f s CPU"
                                  % (n, used, wall, cpu_s),
                        "per_core_value": aln / cpu_s}
        if not a.no_full_loop:
            cpu_baseline["full_loop"] = cpu_full_loop(s, cores)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = hb.Context(local)
    # a real (non-default) stream: the library launches on it and the CUDA events below time it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    reads = s.reads_batch()
    S_tot, R_tot = int(s.locus_sample_off[-1]), int(s.n_reads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident path ------------------------------------------------------------------------
    handle = ctx.upload_genotype(s.batch, reads)
    d_read_ll = torch.zeros(int(s.read_ll_size), dtype=torch.float64, device=dev)
    d_seed = torch.zeros(R_tot, dtype=torch.int32, device=dev)
    d_post = torch.zeros(int(s.post_size), dtype=torch.float64, device=dev)
    d_sll = torch.zeros(S_tot, dtype=torch.float64, device=dev)
    d_best = torch.zeros(S_tot * 2, dtype=torch.int32, device=dev)
    d_tot = torch.zeros(s.n_loci, dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(d_best) for _ in range(world)] if (world > 1 and rank == 0) else None
    launches_per_step = [0]

    def step():
        ctx.genotype_dev(handle, d_read_ll.data_ptr(), d_seed.data_ptr(), d_post.data_ptr(), d_sll.data_ptr(),
                         d_best.data_ptr(), d_tot.data_ptr())
        launches_per_step[0] = ctx.traffic()[2]
        if world > 1:   # the one collective of the path: per-locus genotype records to rank 0
            dist.gather(d_best, gathered, dst=0)

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(a.warmup):
        step()
    barrier()
    ctx.enable_timing(True)
    ctx.collect_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.on = True
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    if clocks:
        clocks.on = False
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    k1_ms, rest_ms, n_calls = ctx.collect_timing()
    ctx.enable_timing(False)
    total_aln = sum_over_ranks(float(n_aln))
    value = total_aln * a.steps / (ms_total / 1e3)
    checksum = float(d_tot.sum().item())
    ctx.free_genotype(handle)

    # ---- end-to-end path: host buffers through the C-ABI ------------------------------------------
    e2e = None
    if not a.no_e2e:
        # page-locked host buffers for the results (the inputs are staged through the library's own
        # page-locked arena)
        read_ll = torch.zeros(int(s.read_ll_size), dtype=torch.float64).pin_memory().numpy()
        read_seed = torch.zeros(R_tot, dtype=torch.int32).pin_memory().numpy()
        h_post = torch.zeros(int(s.post_size), dtype=torch.float64).pin_memory().numpy()
        h_sll = torch.zeros(S_tot, dtype=torch.float64).pin_memory().numpy()
        h_best = torch.zeros(2 * S_tot, dtype=torch.int32).pin_memory().numpy()
        h_tot = torch.zeros(s.n_loci, dtype=torch.float64).pin_memory().numpy()

        def host_step():
            return ctx.genotype_host(s.batch, reads, int(s.read_ll_size), R_tot, int(s.post_size), S_tot, s.n_loci,
                                     read_ll=read_ll, read_seed=read_seed, post=h_post, sample_ll=h_sll, best=h_best,
                                     total_ll=h_tot)
        for _ in range(a.warmup):
            out = host_step()
        barrier()
        if clocks:
            clocks.on = True
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = host_step()
            if world > 1:
                dist.gather(torch.from_numpy(out["best"].ravel()).to(dev), gathered, dst=0)
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        if clocks:
            clocks.on = False
        h2d, d2h, _ = ctx.traffic()
        e2e = {"value": total_aln * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * t_e2e / a.steps, "checksum_matches_resident": bool(abs(float(out["total_ll"].sum()) - checksum) < 1e-6 * abs(checksum))}
    if clocks:
        clocks.stop()
    full_loop = None
    if not a.no_full_loop:
        if world == 1:
            full_loop = gpu_full_loop(local, s, 1)
            full_loop["pipelined"] = gpu_full_loop(local, s, 3)
        else:
            from hipstr_b200.sharding import gather_vcf_records
            n_merged = [0]

            def gather(records, t0):
                merged = gather_vcf_records(records, device=dev)
                barrier()
                if merged is not None:
                    n_merged[0] = len(merged)
                return max_over_ranks(time.perf_counter() - t0)
            barrier()
            full_loop = gpu_full_loop(local, s, 3, gather=gather, locus_base=rank * a.loci)
            full_loop["loci"] = world * a.loci
            full_loop["loci_per_s"] = world * a.loci / full_loop["seconds"]
            full_loop["records_on_rank0"] = n_merged[0]
            full_loop["note"] = "every rank genotypes its own loci; VCF records gathered to rank 0 over NCCL inside the timed region; per-rank counters are rank 0's"

    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        k1_avg_ms = k1_ms / max(n_calls, 1)
        traffic, traffic_src, fp64 = None, None, None
        try:   # DRAM bytes of K1 from the committed ncu --set full capture, scaled per alignment
            tr = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            traffic = tr["dram_bytes_per_alignment"] * n_aln
            traffic_src = tr["capture"]
            fp64 = {"dadd_per_alignment": tr["fp64_dadd_thread_ops_per_alignment"],
                    "peak_dadd_per_s": tr["fp64_dadd_peak_thread_ops_per_cycle"] * 1.965e9}
        except (OSError, ValueError, KeyError):
            pass
        achieved = alg_bytes / (k1_avg_ms / 1e3) / 1e9 if k1_avg_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "loci_per_s": world * a.loci * a.steps / (ms_total / 1e3),
            "alignments_per_step": int(total_aln), "gpu_launches": int(launches_per_step[0]) * a.steps,
            "roofline": {"bound": "hbm", "kernel": "k_align (K1)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch_set": int(alg_bytes), "bytes_per_alignment": alg_bytes / max(n_aln, 1),
                         "k1_ms_per_step": k1_avg_ms, "k1_share_of_step": k1_ms / max(k1_ms + rest_ms, 1e-9),
                         "note": "K1 keeps the DP on chip: it is FP64-issue / latency bound, not HBM bound (DESIGN.md)",
                         "fp64": None if not fp64 else {
                             "achieved": fp64["dadd_per_alignment"] * (n_aln / (k1_avg_ms / 1e3)) if k1_avg_ms > 0 else 0.0,
                             "peak": fp64["peak_dadd_per_s"], "unit": "DADD/s",
                             "frac": fp64["dadd_per_alignment"] * (n_aln / (k1_avg_ms / 1e3)) / fp64["peak_dadd_per_s"] if k1_avg_ms > 0 else 0.0,
                             "source": "ncu DADD count per alignment (profiles/k1_traffic.json) x live alignments/s; peak = 64 FP64 lanes x 148 SMs x 1965 MHz"}},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks.summary() if clocks else None,
            "synth_seconds": t_gen, "checksum_total_ll": checksum, "full_loop": full_loop,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
