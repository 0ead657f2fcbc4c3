#!/usr/bin/env python
"""bench.py -- the HipSTR hot path on B200: read x haplotype HMM alignments/s and loci/s of the genotyping loop.

--workload align (default; BASELINE.json configs[1]: 1 000 synthetic loci, 100 samples x 30 reads, 8 candidate alleles,
150-bp reads, SURVEY.md 8d generator).  One "step" = one pass of the hot path over the batch: K1a stutter tables + K1b
wavefront DP of every pooled read against every candidate haplotype, K2 pool->read scatter + mate merge, K3 genotype
posteriors.  Under torchrun every rank owns its own 1 000 loci (weak scaling: loci are independent, no data-path
collective) and the per-locus genotype records are gathered to rank 0 over NCCL at the end of every step.
  value     alignments/s with the batch resident in HBM (hipstr_genotype_batch_dev)
  e2e       the same through the host-buffer C-ABI call (hipstr_genotype_batch_host): host flattening, H2D of the inputs,
            K1-K3, D2H of read LLs / posteriors / genotypes inside the timed region
  roofline  the binding roof of K1 is FP64 issue (measured DADD / flank-cell rates, profiles/fp64_peak.json); the HBM
            figure the contract asks for is carried as `hbm` beside it; gcups = DP cell updates/s
  full_loop loci/s of the whole seam-B1 path -- constructor from reads, genotype() with allele discovery / pruning / flank
            assembly rounds, write_vcf_record -- through the C++ multi-GPU driver (hipstr_multi_*): ONE shared locus list,
            windows dealt dynamically (to the pipelines of one GPU at N=1; to all ranks through a counter in the rendezvous
            store at N>1, strong scaling), records gathered to rank 0; at N=1 the records of a sample of loci are compared
            with the unmodified reference SeqStutterGenotyper inside the run (a mismatch fails the run)
--workload loop    the loop alone as the metric (loci/s); --loci / --alleles choose the list (configs[2]: 10000 x 16)
--workload cfg4_em K4, the EM stutter learner, on configs[3] (500 samples x 5 reads, 32 alleles) beside the reference's
                   EMStutterGenotyper::train on the host cores
--workload sweep   configs[4]: alleles 2-64 x read length 75-250, alignments/s per point beside the reference

--impl reference times the reference's own CPU code (oracle/_ref/libhipstr_ref.so compiled from the unmodified sources;
the C++ restatement in oracle/ if that library is absent) on the host cores, persistent worker pool, >= 2 loci per core.
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
    write_vcf_record) for locus l of the inherited synthetic batch; returns (seconds, genotype() succeeded, record)."""
    from ref_genotyper import LocusReads, RefGenotyper
    rd = LocusReads(_CPU["loop_synth"], l)
    t0 = time.perf_counter()
    g = RefGenotyper(rd, reassemble_flanks=True)
    ok = g.initialized and g.genotype(1000, 4, 0.01)
    text = g.vcf() if ok else None
    dt = time.perf_counter() - t0
    g.close()
    return dt, bool(ok), text


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuPool:
    """A persistent pool of forked workers (created BEFORE CUDA is initialised) over the inherited synthetic batch."""

    def __init__(self, synth, cores, loop_synth=None):
        lib, prefix, kind = _cpu_lib()
        _CPU.update(synth=synth, lib=lib, prefix=prefix, loop_synth=loop_synth or synth)
        self.synth, self.cores, self.kind = synth, cores, kind
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None

    def align_pass(self, loci_per_core):
        """One bounded CPU sample: every worker aligns `loci_per_core` loci (+ posteriors).  Returns (alignments, wall s,
        CPU s, loci, workers)."""
        s = self.synth
        n = min(s.n_loci, self.cores * loci_per_core)
        cores = min(self.cores, n)
        bounds = [(i * n // cores, (i + 1) * n // cores) for i in range(cores)]
        aln = int(((s.pool_seed[:s.locus_pool_off[n]] >= 0).astype(np.int64) *
                   np.repeat(s.n_haps[:n], np.diff(s.locus_pool_off[:n + 1]))).sum())
        t0 = time.perf_counter()
        secs = self.pool.map(_cpu_worker, bounds) if self.pool else [_cpu_worker(b) for b in bounds]
        return aln, time.perf_counter() - t0, float(sum(secs)), n, cores

    def full_loop(self, loci):
        """The unmodified reference SeqStutterGenotyper on the given loci of the loop list; keeps their records."""
        import checkers
        if checkers.ref() is None:
            return None
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_loop_worker, loci) if self.pool else [_cpu_loop_worker(l) for l in loci]
        wall = time.perf_counter() - t0
        return {"loci_per_s": len(loci) / wall, "cores": min(self.cores, len(loci)),
                "seconds_per_locus_per_core": float(np.mean([r[0] for r in res])),
                "sample": "%d loci spread over the list, %d per core: constructor + genotype(flank assembly on) + write_vcf_record, %.1f s wall"
                          % (len(loci), -(-len(loci) // self.cores), wall),
                "_records": {int(l): (r[1], r[2]) for l, r in zip(loci, res)}}

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


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
    cells = 0
    pool_len = np.diff(s.pool_seq_off).astype(np.int64)
    for l in range(L):
        H, hlen = int(s.n_haps[l]), 0
        for k in range(lbo[l], lbo[l + 1]):
            lens = opt_len[blk_off[k]:blk_off[k + 1]]
            hlen += int(lens.mean())
            if period[k] > 0:
                ndel = np.minimum(6, lens // period[k])
                rep_bytes += int((13 * 8 + (ndel + 1) * 4 * lens).sum())
        hap_bytes += H * (2 * hlen + 64)
        p0, p1 = int(s.locus_pool_off[l]), int(s.locus_pool_off[l + 1])
        seeded = s.pool_seed[p0:p1] >= 0
        cells += int(((pool_len[p0:p1][seeded] - 1) * hlen).sum()) * H    # (L_r - 1) * L_h per alignment
    n_aln = hb.load().hipstr_batch_num_alignments(C.byref(b))
    return read_bytes + hap_bytes + rep_bytes + 8 * n_aln, n_aln, cells


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


def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except (OSError, ValueError):
        return None


# ---------------------------------------------------------------------------------------------
# The genotyping loop through the C++ multi-GPU driver
# ---------------------------------------------------------------------------------------------
def loop_vcf_loci(synth):
    from hipstr_b200.capi import Genotyper
    L = synth.n_loci
    S = int(synth.locus_sample_off[1])
    names = ["S%d" % i for i in range(S)]
    cl = int(synth.view.chrom_len)
    raw = C.string_at(synth.view.chrom_seqs, L * cl)
    period = int(synth.cfg.period) or 4
    return Genotyper.vcf_loci(["chrS"] * L, ["STR"] * L, [synth.view.region_start] * L, [synth.view.region_stop] * L, [period] * L,
                              [raw[l * cl:(l + 1) * cl] for l in range(L)], names * L, names)


def gpu_full_loop(device, segments, pipelines, window, reps, world=1, rank=0, dist=None, torch_dev=None, barrier=None,
                  max_over_ranks=None):
    """Seam B1 end to end: hipstr_multi_genotype (create_from_reads -> genotype(flank assembly) -> write_vcf per window).
    One shared locus list (handed over in `segments`, see main); at N>1 the ranks pull windows from one counter in the
    rendezvous store (dynamic dealing, strong scaling) and the finished VCF records are gathered to rank 0 over NCCL inside
    the timed region."""
    from hipstr_b200.capi import MultiGenotyper
    from hipstr_b200.sharding import StoreDealer, gather_vcf_records
    L = sum(sg.n_loci for sg in segments)
    base = np.concatenate([[0], np.cumsum([sg.n_loci for sg in segments])]).astype(int)
    vls = [loop_vcf_loci(sg) for sg in segments]   # inputs of write_vcf_record are host buffers the caller owns, like the reads
    read_bytes = sum(2 * int(np.ctypeslib.as_array(sg.view.read_seq_off, shape=(int(sg.n_reads) + 1,))[-1]) for sg in segments)
    m = MultiGenotyper(devices=[device], pipelines=pipelines)
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    best = None
    for rep in range(reps):   # the first pass warms the allocations (device buffers, page-locked block cache)
        if barrier:
            barrier()
        t0 = time.perf_counter()
        mine, oks, st = [], [], None
        for k, (sg, vl) in enumerate(zip(segments, vls)):
            dealer = StoreDealer(store, "hipstr_loop_%d_%d" % (rep, k)) if world > 1 else None
            ok_k, rec = m.genotype_synth(sg, vl, window, next_window=dealer)
            mine += [(int(base[k]) + l, "chrS", r[0], r[1]) for l, r in enumerate(rec) if r is not None]
            oks.append(ok_k)
            sk = m.stats()   # per call: summed over the segments here
            if st is None:
                st = sk
            else:
                for key in ("alignments", "traces", "h2d_bytes", "d2h_bytes", "gpu_launches"):
                    st[key] += sk[key]
                st["stage_seconds"] = {n: round(st["stage_seconds"][n] + v, 4) for n, v in sk["stage_seconds"].items()}
                st["windows_per_worker"] = [x + y for x, y in zip(st["windows_per_worker"], sk["windows_per_worker"])]
        ok = np.concatenate(oks)
        n_merged = len(mine)
        merged = mine
        if world > 1:   # the one collective: finished records to rank 0
            merged = gather_vcf_records(mine, device=torch_dev)
            import torch
            ok_all = torch.from_numpy(ok.astype(np.int32)).to(torch_dev)   # a locus is genotyped by the rank that took its window
            dist.all_reduce(ok_all, op=dist.ReduceOp.MAX)
            ok = ok_all.cpu().numpy().astype(np.uint8)
            barrier()
            n_merged = len(merged) if merged is not None else 0
            dt = max_over_ranks(time.perf_counter() - t0)
        else:
            dt = time.perf_counter() - t0
        res = {"loci_per_s": L / dt, "seconds": dt, "loci": L, "segments": [int(sg.n_loci) for sg in segments], "n_gpus": world,
               "pipelines_per_gpu": pipelines, "window_loci": window,
               "scaling": "strong: one shared locus list, windows dealt dynamically" + (" to the ranks through the rendezvous store" if world > 1 else " to the pipelines"),
               "records_on_rank0": n_merged, "alignments_this_rank": st["alignments"], "traces_this_rank": st["traces"],
               "h2d_bytes_this_rank": st["h2d_bytes"], "d2h_bytes_this_rank": st["d2h_bytes"], "gpu_launches_this_rank": st["gpu_launches"],
               "read_bytes_of_the_list": read_bytes,
               "windows_per_worker_this_rank": st["windows_per_worker"], "stage_seconds_summed_over_windows_this_rank": st["stage_seconds"],
               "host_threads": int(os.environ.get("HIPSTR_HOST_THREADS", host_cores())),
               "what": "hipstr_multi_genotype: create_from_reads + genotype(1000, 4, 0.01, reassemble_flanks) + write_vcf per window, host buffers in, VCF text out",
               "_ok": ok, "_records": merged}
        if best is None or res["loci_per_s"] > best["loci_per_s"]:
            best = res
    m.close()
    return best


def check_against_reference(loop, ref_records):
    """Records of the sampled loci vs the unmodified reference SeqStutterGenotyper; a mismatch fails the run."""
    ours = {r[0]: r[3] for r in loop["_records"]}
    norm = lambda t: t.rstrip("\n").replace(":-0.00:", ":0.00:")   # cephes bdtr rounding noise around p = 1
    bad = []
    for l, (ok, text) in sorted(ref_records.items()):
        if bool(loop["_ok"][l]) != ok or (ok and norm(ours.get(l, "")) != norm(text)):
            bad.append(l)
    if bad:
        raise SystemExit("bench.py: loop records differ from the reference SeqStutterGenotyper at loci %s" % bad[:8])
    return {"loci_checked": len(ref_records), "identical": len(ref_records) - len(bad),
            "reference_genotyped": sum(1 for ok, _ in ref_records.values() if ok)}


# ---------------------------------------------------------------------------------------------
def bp_diffs(s):
    return np.ctypeslib.as_array(s.view.read_bp_diff, shape=(int(s.n_reads),))


def em_inputs(s):
    """configs[3]: the EM learner's inputs from the synthetic reads (num_bps = CIGAR bp differences)."""
    from hipstr_b200.capi import EmBatch
    L = s.n_loci
    motif = np.full(L, int(s.cfg.period) or 4, np.int32)
    ref = np.zeros(L, np.int32)
    eb = EmBatch(L, ptr(s.locus_read_off, c_i32p), ptr(s.locus_sample_off, c_i32p), s.view.read_bp_diff,
                 ptr(s.sample_label, c_i32p), ptr(s.log_p1, c_f64p), ptr(s.log_p2, c_f64p), ptr(motif, c_i32p), ptr(ref, c_i32p),
                 ptr(s.haploid, c_u8p))
    eb._keep = (motif, ref)
    return eb


def _cpu_em_worker(rng):
    l0, l1 = rng
    s, lib, prefix = _CPU["synth"], _CPU["lib"], _CPU["prefix"]
    from hipstr_b200.capi import EmBatch
    L = l1 - l0
    motif = np.full(L, int(s.cfg.period) or 4, np.int32)
    ref = np.zeros(L, np.int32)
    r0 = int(s.locus_read_off[l0])
    lro = (s.locus_read_off[l0:l1 + 1] - r0).astype(np.int32)
    lso = (s.locus_sample_off[l0:l1 + 1] - s.locus_sample_off[l0]).astype(np.int32)
    r1 = int(s.locus_read_off[l1])
    arrs = [np.ascontiguousarray(x[r0:r1]) for x in (bp_diffs(s), s.sample_label, s.log_p1, s.log_p2)]
    hap = np.ascontiguousarray(s.haploid[l0:l1])
    eb = EmBatch(L, ptr(lro, c_i32p), ptr(lso, c_i32p), ptr(arrs[0], c_i32p), ptr(arrs[1], c_i32p), ptr(arrs[2], c_f64p),
                 ptr(arrs[3], c_f64p), ptr(motif, c_i32p), ptr(ref, c_i32p), ptr(hap, c_u8p))
    params, conv, iters, ll = np.zeros(6 * L), np.zeros(L, np.uint8), np.zeros(L, np.int32), np.zeros(L)
    t0 = time.perf_counter()
    fn = getattr(lib, prefix + "em_train")
    fn.restype = C.c_int32
    fn.argtypes = [C.POINTER(EmBatch), C.c_int32, C.c_double, C.c_double, c_f64p, c_u8p, c_i32p, c_f64p]
    st = fn(C.byref(eb), 100, 0.01, 0.001, ptr(params, c_f64p), ptr(conv, c_u8p), ptr(iters, c_i32p), ptr(ll, c_f64p))
    assert st == 0
    return time.perf_counter() - t0, params, conv


def workload_cfg4_em(a, rank, world, local):
    """K4 on configs[3]: loci/s of EMStutterGenotyper-equivalent training, host buffers in, 6 parameters per locus out."""
    cores = host_cores()
    s = hb.Synth(n_loci=a.loci, n_samples=500, reads_per_sample=5, n_alleles=32, read_len=a.read_len, seed=3000 + rank)
    cpu = None
    if rank == 0:
        pool = CpuPool(s, cores)
        n = min(s.n_loci, 2 * cores)
        bounds = [(i * n // cores, (i + 1) * n // cores) for i in range(cores) if (i + 1) * n // cores > i * n // cores]
        t0 = time.perf_counter()
        res = pool.pool.map(_cpu_em_worker, bounds) if pool.pool else [_cpu_em_worker(b) for b in bounds]
        wall = time.perf_counter() - t0
        cpu = {"value": n / wall, "unit": "loci/s", "cores": len(bounds), "kind": pool.kind,
               "sample": "first %d loci, EMStutterGenotyper::train(100, 0.01, 0.001), %d forked workers, %.2f s wall" % (n, len(bounds), wall),
               "_params": np.concatenate([r[1] for r in res]), "_conv": np.concatenate([r[2] for r in res]), "_n": n}
        pool.close()
    import torch
    torch.cuda.set_device(local)
    ctx = hb.Context(local)
    eb = em_inputs(s)
    L = s.n_loci
    params, conv, iters, ll = np.zeros(6 * L), np.zeros(L, np.uint8), np.zeros(L, np.int32), np.zeros(L)

    def step():
        st = ctx.lib.hipstr_em_train_host(ctx.h, C.byref(eb), 100, 0.01, 0.001, ptr(params, c_f64p), ptr(conv, c_u8p), ptr(iters, c_i32p),
                                          ptr(ll, c_f64p))
        assert st == 0
    clocks = ClockSampler(local)
    for _ in range(a.warmup):
        step()
    clocks.on = True
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    clocks.on = False
    clocks.stop()
    check = None
    if cpu:
        n = cpu.pop("_n")
        err = float(np.abs(params[:6 * n] - cpu.pop("_params")).max())
        same_conv = bool((conv[:n] == cpu.pop("_conv")).all())
        if err > 1e-9 or not same_conv:
            raise SystemExit("bench.py: EM parameters differ from the reference (max |diff| %.3g, convergence flags equal: %s)" % (err, same_conv))
        check = {"loci_checked": n, "max_abs_param_diff": err, "convergence_flags_equal": same_conv}
    v = L * a.steps / dt
    emit({"metric": "EM stutter-model training, loci/sec", "value": v, "unit": "loci/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
          "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "%d synthetic loci, 500 samples x 5 reads, 32 alleles (BASELINE.json configs[3]): hipstr_em_train_host, one 512-thread CTA per locus" % L},
          "e2e": {"value": v, "unit": "loci/s", "h2d_bytes_per_step": int(ctx.traffic()[0]), "d2h_bytes_per_step": int(ctx.traffic()[1])},
          "gpu_launches": int(ctx.traffic()[2]) * a.steps, "cpu_baseline": cpu, "reference_check": check,
          "mean_iterations": float(iters.mean()), "converged": int(conv.sum()), "clocks": clocks.summary()})


def workload_sweep(a, rank, world, local):
    """configs[4]: alleles 2-64 x read length 75-250 (trim off so that the read length bites), alignments/s of K1
    resident in HBM per point, beside the reference on the host cores (bounded sample)."""
    import torch
    cores = host_cores()
    points = [(al, rl) for al in (2, 4, 8, 16, 32, 64) for rl in (75, 100, 150, 200, 250)]
    loci = max(8, a.loci // 20)
    synths = {p: hb.Synth(n_loci=loci, n_samples=100, reads_per_sample=30, n_alleles=p[0], read_len=p[1], seed=4000, trim=0) for p in points}
    cpu = {}
    for p in points:   # before CUDA is initialised: workers are forked
        pool = CpuPool(synths[p], cores)
        aln, wall, cpu_s, n, used = pool.align_pass(1)
        cpu[p] = aln / wall
        pool.close()
    torch.cuda.set_device(local)
    ctx = hb.Context(local)
    table = []
    for p in points:
        s = synths[p]
        _, n_aln, cells = algorithmic_bytes(s)
        h = ctx.upload(s.batch)
        out = torch.zeros(s.n_out, dtype=torch.float64, device="cuda:%d" % local)
        ctx.enable_timing(True)
        for _ in range(2):
            ctx.align_dev(h, out.data_ptr())
        ctx.collect_timing()
        for _ in range(a.steps):
            ctx.align_dev(h, out.data_ptr())
        ms = ctx.collect_timing()[0] / a.steps
        ctx.free_batch(h)
        table.append({"alleles": p[0], "read_len": p[1], "alignments": int(n_aln), "alignments_per_s": n_aln / (ms / 1e3),
                      "gcups": cells / (ms / 1e3) / 1e9, "cpu_alignments_per_s": cpu[p], "ratio": n_aln / (ms / 1e3) / cpu[p]})
    best = max(t["alignments_per_s"] for t in table)
    emit({"metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": 2, "ms_per_step": None,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "configs[4] sweep: alleles 2-64 x read length 75-250, %d loci x 100 samples x 30 reads per point, trim off; value = best point" % loci},
          "cpu_baseline": {"value": max(cpu.values()), "unit": UNIT, "cores": cores, "kind": _cpu_lib()[2],
                           "sample": "per point: %d loci, 1 per core, align + posteriors" % min(loci, cores)},
          "sweep": table})


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="align", choices=["align", "loop", "cfg4_em", "sweep"])
    ap.add_argument("--loci", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--reads-per-sample", type=int, default=30)
    ap.add_argument("--alleles", type=int, default=8)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--loop-loci", type=int, default=0, help="loci of the shared list of full_loop (default: --loci at N=1, 1000 per GPU at N>1: enough windows for every pipeline)")
    ap.add_argument("--pipelines", type=int, default=0, help="window pipelines per GPU (default: 4; twice the host threads of a rank, at most 8, when a rank has fewer than 8 threads)")
    ap.add_argument("--window", type=int, default=50)
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
    if a.pipelines <= 0:
        # With few host threads per GPU (8 GPUs on 32 cores) a pipeline is one thread, and it sleeps while its device call runs:
        # twice as many pipelines as threads keep the cores busy (measured on 4 cores per rank: 4 pipelines 970-1 000 loci/s,
        # 6: 1 120, 8: 1 135); with 12-16 threads per GPU four pipelines of 3-4 threads are best
        threads = int(os.environ.get("HIPSTR_HOST_THREADS", host_cores()))
        a.pipelines = 4 if threads >= 8 else max(4, min(8, 2 * threads))
    if a.impl == "ours" and a.workload == "cfg4_em":
        if rank == 0:
            workload_cfg4_em(a, rank, world, local)
        return
    if a.impl == "ours" and a.workload == "sweep":
        if rank == 0:
            workload_sweep(a, rank, world, local)
        return
    workload = "%d synthetic loci, %d samples x %d reads, %d alleles, %d bp reads (BASELINE.json configs[1])" % (
        a.loci, a.samples, a.reads_per_sample, a.alleles, a.read_len)
    config = {"workload": workload, "loci_per_gpu": a.loci, "sharding": "independent loci per rank, NCCL gather of per-locus genotype records per step",
              "l2": "inputs+outputs of a step (~0.6 GB) and the stutter tables (GBs per chunk) exceed the 126 MB L2; no explicit flush"}

    if a.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        per_core = 2
        s = hb.Synth(n_loci=min(a.loci, per_core * cores), n_samples=a.samples, reads_per_sample=a.reads_per_sample,
                     n_alleles=a.alleles, read_len=a.read_len, seed=2000)
        pool = CpuPool(s, cores)
        for _ in range(a.warmup):
            pool.align_pass(per_core)
        tot_aln, tot_t = 0, 0.0
        for _ in range(a.steps):
            aln, wall, _, n, used = pool.align_pass(per_core)
            tot_aln += aln
            tot_t += wall
        pool.close()
        v = tot_aln / tot_t
        sample = "%d loci per step (%d per core, persistent pool of %d forked workers) of the same synthetic workload" % (n, per_core, used)
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": pool.kind, "sample": sample},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}))
        return

    loop_only = a.workload == "loop"
    t_gen = time.time()
    # the loop's shared list: the same loci on every rank (windows are dealt dynamically, any rank may get any window)
    loop_loci = a.loop_loci or (a.loci if (world == 1 or loop_only) else max(2000, 1000 * world))
    # The C-ABI's read offsets are 32-bit (hipstr_locus_reads_t): a list whose read bytes pass ~1.5 GB is handed over in
    # segments, each one hipstr_multi_genotype call with its own dealer (configs[2], 10 000 loci x 3 000 reads x 150 bp = 4.5 GB)
    seg_loci = max(a.window, int(1.5e9 / max(1, a.samples * a.reads_per_sample * a.read_len)) // a.window * a.window)
    seg_sizes = [min(seg_loci, loop_loci - k) for k in range(0, loop_loci, seg_loci)]
    if not loop_only and a.loci > seg_loci:
        raise SystemExit("bench.py: --loci %d passes the 32-bit read offsets of one batch (at most %d loci of this shape)" % (a.loci, seg_loci))
    s = hb.Synth(n_loci=seg_sizes[0] if loop_only else a.loci, n_samples=a.samples, reads_per_sample=a.reads_per_sample,
                 n_alleles=a.alleles, read_len=a.read_len, seed=2000 + (0 if loop_only else rank))
    if (world == 1 or loop_only) and loop_loci == a.loci and len(seg_sizes) == 1:
        loop_segments = [s]
    elif loop_only:
        loop_segments = [s] + [hb.Synth(n_loci=n, n_samples=a.samples, reads_per_sample=a.reads_per_sample, n_alleles=a.alleles,
                                        read_len=a.read_len, seed=2000 + 7919 * k) for k, n in enumerate(seg_sizes) if k > 0]
    elif a.no_full_loop:
        loop_segments = []
    else:
        loop_segments = [hb.Synth(n_loci=n, n_samples=a.samples, reads_per_sample=a.reads_per_sample, n_alleles=a.alleles,
                                  read_len=a.read_len, seed=2000 + 7919 * k) for k, n in enumerate(seg_sizes)]
    s_loop = loop_segments[0] if loop_segments else None   # the reference check samples its loci from the first segment
    t_gen = time.time() - t_gen
    alg_bytes, n_aln, cells = algorithmic_bytes(s)

    cpu_baseline, ref_records = None, None
    if rank == 0 and not a.no_cpu_baseline:   # before CUDA is initialised: workers are forked
        cores = host_cores()
        pool = CpuPool(s, cores, loop_synth=s_loop)
        if not loop_only:
            pool.align_pass(1)   # warm the pool
            aln, wall, cpu_s, n, used = pool.align_pass(2)
            cpu_baseline = {"value": aln / wall, "unit": UNIT, "cores": used, "kind": pool.kind,
                            "sample": "first %d loci of the workload (2 per core), align + posteriors, persistent pool of %d forked workers, %.1f s wall / %.1f s CPU"
                                      % (n, used, wall, cpu_s),
                            "per_core_value": aln / cpu_s}
        else:
            cpu_baseline = {"unit": "loci/s", "cores": cores, "kind": pool.kind}
        if not a.no_full_loop and s_loop is not None:
            n_check = min(s_loop.n_loci, max(32, 2 * cores))
            sample_loci = sorted(set(int(x) for x in np.linspace(0, s_loop.n_loci - 1, n_check)))
            fl = pool.full_loop(sample_loci)
            if fl:
                ref_records = fl.pop("_records")
                cpu_baseline["full_loop"] = fl
                if loop_only:
                    cpu_baseline.update(value=fl["loci_per_s"], sample=fl["sample"])
        pool.close()

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

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

    clocks = ClockSampler(local) if rank == 0 else None

    def run_loop():
        if clocks:
            clocks.on = True
        fl = gpu_full_loop(local, loop_segments, a.pipelines, a.window, 3, world=world, rank=rank, dist=dist, torch_dev=dev,
                           barrier=barrier, max_over_ranks=max_over_ranks)
        if clocks:
            clocks.on = False
        if rank == 0 and ref_records:
            fl["reference_check"] = check_against_reference(fl, ref_records)
        fl.pop("_ok")
        fl.pop("_records")
        return fl

    if loop_only:
        fl = run_loop()
        if clocks:
            clocks.stop()
        if rank == 0:
            emit({"metric": "STR loci genotyped/sec (SeqStutterGenotyper loop: constructor, genotype(), write_vcf_record)", "value": fl["loci_per_s"],
                  "unit": "loci/s", "n_gpus": world, "steps": 1, "warmup": 2, "ms_per_step": 1e3 * fl["seconds"], "higher_is_better": True,
                  "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": "%d synthetic loci, %d samples x %d reads, %d alleles, %d bp reads; one shared list, windows of %d loci dealt dynamically"
                                         % (loop_loci, a.samples, a.reads_per_sample, a.alleles, a.read_len, a.window)},
                  "e2e": {"value": fl["loci_per_s"], "unit": "loci/s", "h2d_bytes_per_step": fl["h2d_bytes_this_rank"],
                          "d2h_bytes_per_step": fl["d2h_bytes_this_rank"],
                          "note": "the loop's inputs are host buffers and its output is VCF text: value is already end to end; bytes and launches are rank 0's share"},
                  "gpu_launches": fl["gpu_launches_this_rank"], "cpu_baseline": cpu_baseline, "clocks": clocks.summary() if clocks else None, "full_loop": fl})
        if world > 1:
            dist.destroy_process_group()
        return

    ctx = hb.Context(local)
    # a real (non-default) stream: the library launches on it and the CUDA events below time it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    reads = s.reads_batch()
    S_tot, R_tot = int(s.locus_sample_off[-1]), int(s.n_reads)

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
    total_cells = sum_over_ranks(float(cells))
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
    ctx.close()
    full_loop = run_loop() if (not a.no_full_loop and s_loop is not None) else None
    if clocks:
        clocks.stop()

    if rank == 0:
        peaks = load_json("MEASURED_PEAKS.json")
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback"
        hbm_peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        k1_avg_ms = k1_ms / max(n_calls, 1)
        k1_s = k1_avg_ms / 1e3
        tr = load_json("profiles", "k1_traffic.json") or {}
        fp = load_json("profiles", "fp64_peak.json") or {}
        aln_rank = float(n_aln)
        hbm_achieved = alg_bytes / k1_s / 1e9 if k1_s > 0 else 0.0
        traffic = tr.get("dram_bytes_per_alignment", 0) * aln_rank if tr else None
        # the binding roof: FP64 issue.  One flank-cell update = 9 DADD + 4 double max; the chip-wide rate of that
        # mix was measured on this pool (tools/fp64_peak.cu); every alignment updates (L_r - 1) x L_h cells.
        cell_peak = fp.get("flank_cells_per_s")
        gcups = cells / k1_s / 1e9 if k1_s > 0 else 0.0
        roofline = {"bound": "fp64_issue", "kernel": "K1 = k_stutter (K1a) + k_align (K1b)", "achieved": gcups,
                    "peak": cell_peak / 1e9 if cell_peak else None, "unit": "G DP-cell updates/s",
                    "frac": gcups / (cell_peak / 1e9) if cell_peak else None, "traffic": traffic,
                    "traffic_source": tr.get("capture"),
                    "peak_source": "measured on this pool's B200: tools/fp64_peak.cu, 9 DADD + 4 max per cell with 8 independent chains per thread (profiles/fp64_peak.json); every repeat-block column is counted as ONE cell update although it evaluates 13 artifact sizes",
                    "cells_per_launch_set": int(cells), "k1_ms_per_step": k1_avg_ms, "k1_share_of_step": k1_ms / max(k1_ms + rest_ms, 1e-9),
                    "dadd_per_s_peak": fp.get("dadd_per_s"),
                    "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                            "peak_source": peak_src, "algorithmic_bytes_per_launch_set": int(alg_bytes),
                            "bytes_per_alignment": alg_bytes / max(n_aln, 1),
                            "note": "algorithmic bytes (SURVEY 8d, ~40 B per alignment); K1 keeps the DP on chip, so it is FP64-issue / latency bound, not HBM bound; the stutter tables K1a hands to K1b add 24 KB per alignment of real DRAM traffic (see traffic)"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "gcups": total_cells * a.steps / (ms_total / 1e3) / 1e9,
            "k123_passes_per_s_in_loci": world * a.loci * a.steps / (ms_total / 1e3),
            "alignments_per_step": int(total_aln), "gpu_launches": int(launches_per_step[0]) * a.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks.summary() if clocks else None,
            "synth_seconds": t_gen, "checksum_total_ll": checksum, "full_loop": full_loop,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
