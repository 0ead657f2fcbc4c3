"""From BAM files to VCF records, a window of regions at a time: the stages of the reference's per-region driver
(BamProcessor::process_regions, src/bam_processor.cpp:521-617 -> SNPBamProcessor::process_reads, src/snp_bam_processor.cpp:36-118 ->
GenotyperBamProcessor::analyze_reads_and_phasing, src/genotyper_bam_processor.cpp:160-289) chained over the C-ABI:

    per region   hipstr_bam_reader_fetch -> hipstr_filter_reads (+ PCR duplicates)                      host
    per window   [hipstr_extract_cigar + hipstr_em_train_host (K4) when no default stutter model]        one launch
                 hipstr_left_align_reads_host (K6)                                                      one launch per NW round
                 hipstr_genotyper_create_from_reads -> genotype (K1 K2 K3 K5 in lockstep rounds) -> write_vcf (K3b K5)

This module is the orchestration a caller writes; every stage is a call into libhipstr_b200.so.  Without a phased SNP VCF the
phasing log-likelihoods are 0 for every read, as in the reference (snp_bam_processor.cpp:100-110); with SNP sets they come
from Context.snp_phasing (K7) on FilteredReads.view().
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import BamReader, Genotyper, LeftAligned, SnpPhasing, SnpVcf, c_i32p, make_em_batch, make_locus_reads, ptr

DEFAULT_STUTTER = (0.95, 0.05, 0.05, 0.95, 0.01, 0.01)   # hipstr_main.cpp:343


class Options:
    """The knobs of GenotyperBamProcessor / BamProcessor that the stages read, with the reference's defaults."""
    max_str_length = 100            # MAX_STR_LENGTH
    min_total_reads = 100           # MIN_TOTAL_READS
    max_total_haplotypes = 1000
    max_flank_haplotypes = 4
    min_flank_freq = 0.01
    max_em_iter, abs_ll_converge, frac_ll_converge = 100, 0.01, 0.001
    def_stutter_model = None        # six parameters, or None = learn the model with the EM stutter genotyper
    recalc_stutter_model = False
    haploid_chroms = ()
    snp_vcf = None                  # path of a phased SNP VCF (bgzipped or plain): phasing log-likelihoods from K7
    max_mate_dist = 1000            # MAX_MATE_DIST
    skip_padding = 15               # SNPBamProcessor::SKIP_PADDING
    filter = None                   # dict of hipstr_filter_options_t overrides
    host_threads = 0                # 0 = HIPSTR_HOST_THREADS / all cores
    bams_from_10x = False           # phasing from the reads' HP tags (native driver only)
    ref_vcf = None                  # path of a reference panel of STR genotypes (--ref-vcf; native driver only)

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(type(self), k):
                raise TypeError("unknown option " + k)
            setattr(self, k, v)


def read_fasta(path):
    seqs, name = {}, None
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                name = line[1:].split()[0]
                seqs[name] = []
            elif name is not None:
                seqs[name].append(line.strip())
    return {k: "".join(v) for k, v in seqs.items()}


def read_regions(path):
    """readRegions + orderRegions (src/region.cpp:14-55): CHROM START(1-based) STOP PERIOD NCOPIES [NAME]."""
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if len(t) < 5:
                raise ValueError("Improperly formatted region file: " + line)
            out.append((t[0], int(t[1]) - 1, int(t[2]), int(t[3]), t[5] if len(t) > 5 else ""))
    return sorted(out, key=lambda r: (r[0], r[1], r[2]))


def _extract_cigar(lib, types, lens, start, region_start, region_end):
    bp = C.c_int32()
    lens = np.ascontiguousarray(lens, np.int32)
    ok = lib.hipstr_extract_cigar(types.encode(), ptr(lens, c_i32p), len(types), start, region_start, region_end, C.byref(bp))
    return bp.value if ok else None


STATUS = ("genotyped", "too_long", "near_contig_end", "too_few_reads", "too_many_reads", "em_failed", "genotype_failed", "unknown_chromosome")


def process_regions(ctx, bam_paths, chrom_seqs, regions, options=None, vcf_options=None):
    """hipstr_process_regions (hipstr_b200/host/region_driver.cpp): the whole chain in native code, one call per window.
    Returns (records, summary) like process_regions_staged."""
    opt = options or Options()
    lib = capi.load()
    po = capi.PipelineOptions()
    lib.hipstr_pipeline_default_options(C.byref(po))
    for k, v in (opt.filter or {}).items():
        setattr(po.filter, k, v)
    po.filter.max_mate_dist = opt.max_mate_dist
    for k in ("max_str_length", "min_total_reads", "max_total_haplotypes", "max_flank_haplotypes", "min_flank_freq", "max_em_iter",
              "abs_ll_converge", "frac_ll_converge", "skip_padding"):
        setattr(po, k, getattr(opt, k))
    po.recalc_stutter_model = int(bool(opt.recalc_stutter_model))
    if opt.def_stutter_model is not None:
        po.use_def_stutter_model = 1
        po.def_stutter_model = (C.c_double * 6)(*opt.def_stutter_model)
    mk = lambda xs: (C.c_char_p * max(len(xs), 1))(*[x if isinstance(x, bytes) else x.encode() for x in xs])
    hap = mk(list(opt.haploid_chroms))
    po.n_haploid_chroms, po.haploid_chroms = len(opt.haploid_chroms), hap
    po.host_threads = opt.host_threads
    po.bams_from_10x = int(bool(opt.bams_from_10x))
    panel = capi.StrVcf(opt.ref_vcf) if opt.ref_vcf else None
    po.ref_vcf = panel.h if panel else None
    vo = capi.VcfOptions()
    lib.hipstr_vcf_default_options(C.byref(vo))
    for k, v in (vcf_options or {}).items():
        setattr(vo, k, v)
    snp_vcf = SnpVcf(opt.snp_vcf) if opt.snp_vcf else None
    names = list(chrom_seqs)
    starts, stops, periods = (np.array([r[k] for r in regions], np.int32) for k in (1, 2, 3))
    h = C.c_void_p()
    st = lib.hipstr_process_regions(ctx.h, len(bam_paths), mk(bam_paths), snp_vcf.h if snp_vcf else None, len(names), mk(names),
                                    mk([chrom_seqs[n] for n in names]), len(regions), mk([r[0] for r in regions]), ptr(starts, c_i32p),
                                    ptr(stops, c_i32p), ptr(periods, c_i32p), mk([r[4] for r in regions]), C.byref(po), C.byref(vo), C.byref(h))
    if st != 0:
        raise capi.HipstrError(st, "process_regions: " + lib.hipstr_process_regions_last_error().decode())
    summary = {k: 0 for k in STATUS}
    records = []
    for i, r in enumerate(regions):
        pos = C.c_int32()
        status = lib.hipstr_region_results_status(h, i, C.byref(pos), None)
        summary[STATUS[status]] += 1
        if status == 0:
            records.append((r[0], pos.value, lib.hipstr_region_results_record(h, i).decode()))
    sec, cnt = np.zeros(8), np.zeros(4, np.int64)
    lib.hipstr_region_results_timing(h, ptr(sec, capi.c_f64p), ptr(cnt, capi.c_i64p))
    summary["seconds"] = dict(zip(("ingest", "phasing", "stutter", "left_align", "genotype", "records", "phasing_pack", "phasing_k7_call"), (float(x) for x in sec)))
    summary["alignments_read"], summary["reads_kept"], summary["phased_reads"], summary["left_align_failed"] = (int(x) for x in cnt)
    gsec, gst = np.zeros(9), np.zeros(3, np.int64)
    lib.hipstr_region_results_genotyper_timing(h, ptr(gsec, capi.c_f64p), ptr(gst, capi.c_i64p))
    summary["genotyper_seconds"] = dict(zip(("construct", "decide", "trace_device", "trace_host", "align", "posteriors", "vcf", "align_pack",
                                             "align_unpack"), (round(float(x), 4) for x in gsec)))
    summary["alignments"], summary["traces"], summary["rounds"] = (int(x) for x in gst)
    summary["samples"] = lib.hipstr_region_results_samples(h).decode().splitlines()
    lib.hipstr_region_results_free(h)
    return records, summary


def process_regions_staged(ctx, bam_paths, chrom_seqs, regions, options=None, vcf_options=None):
    """The same chain stage by stage from Python (what a caller writes against the individual entry points).
    Genotypes `regions` [(chrom, start, stop, period, name)] from the BAM files; returns (records, summary) with
    records = [(chrom, pos, VCF record text)] in region order for the loci that were genotyped."""
    opt = options or Options()
    lib = capi.load()
    lib.hipstr_extract_cigar.restype = C.c_int32
    lib.hipstr_extract_cigar.argtypes = [C.c_char_p, c_i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i32p]
    reader = BamReader(bam_paths)
    rg_map, all_samples = {}, set()
    for path, rg_id, sample, library in reader.read_groups():
        if sample is None or library is None:
            raise ValueError("RG in BAM header is lacking the SM or LB tag")
        rg_map[path + rg_id] = (sample, library)
        all_samples.add(sample)
    out_samples = sorted(all_samples)     # samples_to_genotype_ (genotyper_bam_processor.h:181-190)
    summary = dict(too_long=0, near_contig_end=0, too_few_reads=0, too_many_reads=0, em_failed=0, genotype_failed=0, genotyped=0)

    # ---- per region: ingestion (host) ----------------------------------------------------------------------------
    loci = []
    filter_kw = dict(opt.filter or {})
    filter_kw.setdefault("max_mate_dist", opt.max_mate_dist)
    max_mate_dist = opt.max_mate_dist
    snp_vcf = SnpVcf(opt.snp_vcf) if opt.snp_vcf else None
    vcf_index = {name: i for i, name in enumerate(snp_vcf.samples)} if snp_vcf else {}
    for chrom, start, stop, period, name in regions:
        if stop - start > opt.max_str_length:
            summary["too_long"] += 1
            continue
        seq = chrom_seqs[chrom]
        if start < 50 or stop + 50 >= len(seq):
            summary["near_contig_end"] += 1
            continue
        recs = reader.fetch(chrom, 0 if start < max_mate_dist else start - max_mate_dist, stop + max_mate_dist)
        kept = recs.filter(seq, [(start, stop)], rg_map, **filter_kw)
        counts = kept.counts()
        names = kept.entry_names()                   # (builds a view of its own: taken before the one used below)
        view = kept.view()
        b = view.reads
        n_entries = b.n_entries
        if n_entries < opt.min_total_reads:          # analyze_reads_and_phasing: total_reads < MIN_TOTAL_READS
            summary["too_few_reads"] += 1
            continue
        if counts["too_many_reads"]:
            summary["too_many_reads"] += 1
            continue
        samples = [view.sample_names[i].decode() for i in range(view.n_samples)]
        entry_off = np.ctypeslib.as_array(view.sample_entry_off, shape=(view.n_samples + 1,)).copy()
        aln_off = np.ctypeslib.as_array(b.entry_aln_off, shape=(n_entries + 1,))
        seq_off = np.ctypeslib.as_array(b.aln_seq_off, shape=(b.n_alns + 1,))
        cig_off = np.ctypeslib.as_array(b.aln_cigar_off, shape=(b.n_alns + 1,))
        pos = np.ctypeslib.as_array(b.aln_pos, shape=(b.n_alns,))
        end = np.ctypeslib.as_array(b.aln_end, shape=(b.n_alns,))
        flag = np.ctypeslib.as_array(view.aln_flag, shape=(b.n_alns,))
        bases = C.string_at(b.bases, int(seq_off[-1])).decode("latin-1")
        quals = C.string_at(b.quals, int(seq_off[-1])).decode("latin-1")
        ctype = C.string_at(b.cigar_type, int(cig_off[-1])).decode()
        clen = np.ctypeslib.as_array(b.cigar_len, shape=(max(int(cig_off[-1]), 1),))
        passes = C.string_at(view.entry_passes, n_entries)
        reads, labels, rev, use, name_ids, ids = [], [], [], [], [], {}
        for s in range(view.n_samples):
            for e in range(int(entry_off[s]), int(entry_off[s + 1])):
                a = int(aln_off[e])                   # the STR read; its mate (if any) only matters for phasing
                cig = [(ctype[c], int(clen[c])) for c in range(int(cig_off[a]), int(cig_off[a + 1]))]
                reads.append((int(pos[a]), int(end[a]), bases[seq_off[a]:seq_off[a + 1]], quals[seq_off[a]:seq_off[a + 1]], cig))
                labels.append(s)
                rev.append(1 if flag[a] & 0x10 else 0)
                use.append(1 if passes[e:e + 1] == b"1" else 0)
                name_ids.append(ids.setdefault(names[e], len(ids)))
        L = dict(chrom=chrom, start=start, stop=stop, period=period, name=name, seq=seq, samples=samples, reads=reads,
                 labels=labels, rev=rev, use=use, name_ids=name_ids, haploid=1 if chrom in opt.haploid_chroms else 0)
        sets = snp_vcf.region_sets(chrom, start - max_mate_dist if start > max_mate_dist else 1, stop + max_mate_dist, [(start, stop)],
                                   opt.skip_padding) if snp_vcf else None
        if sets is not None:   # copies of the view (it dies with `kept`): this locus' share of the window's K7 batch
            entry_set = np.repeat([vcf_index.get(smp, -1) for smp in samples], np.diff(entry_off))
            L["phasing"] = dict(entry_aln_off=aln_off.copy(), entry_set=entry_set, aln_pos=pos.copy(), aln_end=end.copy(),
                                aln_seq_off=seq_off.copy(), bases=bases.encode("latin-1"), quals=quals.encode("latin-1"),
                                aln_cigar_off=cig_off.copy(), cigar_type=ctype.encode(), cigar_len=clen[:int(cig_off[-1])].copy(), sets=sets)
        loci.append(L)

    # ---- phasing log-likelihoods of every read of the window in ONE K7 launch ------------------------------------------
    for L in loci:
        L["log_p1"], L["log_p2"] = np.zeros(len(L["reads"])), np.zeros(len(L["reads"]))
    parts = [L["phasing"] for L in loci if "phasing" in L]
    if parts:
        aln_shift = np.cumsum([0] + [len(p["aln_pos"]) for p in parts])
        seq_shift = np.cumsum([0] + [len(p["bases"]) for p in parts])
        cig_shift = np.cumsum([0] + [len(p["cigar_type"]) for p in parts])
        set_shift = np.cumsum([0] + [len(p["sets"][0]) - 1 for p in parts])
        snp_shift = np.cumsum([0] + [len(p["sets"][1]) for p in parts])
        offs = lambda key, shift: np.concatenate([np.asarray(p[key])[:-1] + s for p, s in zip(parts, shift)] + [[shift[-1]]])
        batch = SnpPhasing.from_arrays(
            offs("entry_aln_off", aln_shift), np.concatenate([np.where(p["entry_set"] >= 0, p["entry_set"] + s, -1) for p, s in zip(parts, set_shift)]),
            np.concatenate([p["aln_pos"] for p in parts]), np.concatenate([p["aln_end"] for p in parts]), offs("aln_seq_off", seq_shift),
            b"".join(p["bases"] for p in parts), b"".join(p["quals"] for p in parts), offs("aln_cigar_off", cig_shift),
            b"".join(p["cigar_type"] for p in parts), np.concatenate([p["cigar_len"] for p in parts]),
            np.concatenate([np.asarray(p["sets"][0])[:-1] + s for p, s in zip(parts, snp_shift)] + [[snp_shift[-1]]]),
            np.concatenate([p["sets"][1] for p in parts]), b"".join(p["sets"][2] for p in parts), b"".join(p["sets"][3] for p in parts))
        p1, p2, counts = ctx.snp_phasing(batch)
        at = 0
        for L in loci:
            if "phasing" in L:
                n_e = len(L["reads"])
                L["log_p1"], L["log_p2"] = p1[at:at + n_e].copy(), p2[at:at + n_e].copy()
                at += n_e
        summary["snp_matches"], summary["snp_mismatches"] = int(counts[:, :2].sum()), int(counts[:, 2].sum())
        summary["phased_reads"] = int((p1 != p2).sum())

    # ---- stutter models: the default, or one EM fit per locus, all loci in one K4 call --------------------------------
    if opt.def_stutter_model is not None:
        for L in loci:
            L["stutter"] = tuple(opt.def_stutter_model)
    elif loci:
        lro, lso, num_bps, labels, em_p1, em_p2 = [0], [0], [], [], [], []
        for L in loci:
            informative = 0
            per_sample = [[] for _ in L["samples"]]
            for r, ((start, _, _, _, cig), s) in enumerate(zip(L["reads"], L["labels"])):
                if informative > 10000 and not per_sample[s]:      # MAX_INF_READS is checked between samples
                    continue
                bp = _extract_cigar(lib, "".join(t for t, _ in cig), [n for _, n in cig], start, L["start"] - L["period"], L["stop"] + L["period"])
                if bp is None or bp < -(L["stop"] - L["start"] + 1):
                    continue
                per_sample[s].append((bp, L["log_p1"][r], L["log_p2"][r]))
                informative += 1
            L["informative"] = informative
            for s, bps in enumerate(per_sample):
                num_bps += [x[0] for x in bps]
                em_p1 += [x[1] for x in bps]
                em_p2 += [x[2] for x in bps]
                labels += [s] * len(bps)
            lro.append(len(num_bps))
            lso.append(lso[-1] + len(L["samples"]))
        batch = make_em_batch(lro, lso, num_bps, labels, em_p1, em_p2, [L["period"] for L in loci], [0] * len(loci), [L["haploid"] for L in loci])
        params, converged, _, _ = ctx.em_train(batch, opt.max_em_iter, opt.abs_ll_converge, opt.frac_ll_converge)
        trained = []
        for l, L in enumerate(loci):
            if L["informative"] < opt.min_total_reads:
                summary["too_few_reads"] += 1
            elif not converged[l]:
                summary["em_failed"] += 1
            else:
                L["stutter"] = tuple(float(x) for x in params[l])
                trained.append(L)
        loci = trained
    if not loci:
        return [], summary

    # ---- left alignment (K6) of all loci, then the lockstep genotyper over the window -----------------------------------
    n = len(loci)
    read_off = np.cumsum([0] + [len(L["reads"]) for L in loci]).astype(np.int32)
    sample_off = np.cumsum([0] + [len(L["samples"]) for L in loci]).astype(np.int32)
    flat = lambda key: [x for L in loci for x in L[key]]
    raw = make_locus_reads(read_off, sample_off, flat("reads"), flat("labels"), flat("name_ids"), np.concatenate([L["log_p1"] for L in loci]),
                           np.concatenate([L["log_p2"] for L in loci]), [L["haploid"] for L in loci], flat("rev"), flat("use"))
    seqs = [L["seq"] for L in loci]
    aligned = LeftAligned(ctx, n, raw, seqs, [L["start"] - 40 if L["start"] > 40 else 1 for L in loci], [L["stop"] + 40 for L in loci])
    g = Genotyper.from_reads(ctx, aligned.view, n, [L["start"] for L in loci], [L["stop"] for L in loci], [L["period"] for L in loci], seqs,
                             stutter=np.array([L["stutter"] for L in loci]))
    ok = g.genotype(opt.max_total_haplotypes, opt.max_flank_haplotypes, opt.min_flank_freq, True)
    if opt.recalc_stutter_model:
        ok = g.recompute_stutter_models(opt.max_total_haplotypes, opt.max_flank_haplotypes, opt.min_flank_freq, opt.max_em_iter,
                                        opt.abs_ll_converge, opt.frac_ll_converge)
    vloci = g.vcf_loci([L["chrom"] for L in loci], [L["name"] for L in loci], [L["start"] for L in loci], [L["stop"] for L in loci],
                       [L["period"] for L in loci], seqs, [s for L in loci for s in L["samples"]], out_samples)
    records = g.write_vcf(vloci, **(vcf_options or {}))
    out = []
    for l, L in enumerate(loci):
        if ok[l] and records[l] is not None:
            out.append((L["chrom"], records[l][0], records[l][1]))
            summary["genotyped"] += 1
        else:
            summary["genotype_failed"] += 1
    summary["left_align_failed"] = int(aligned.failed)
    g.close()
    aligned.close()
    return out, summary


def write_vcf_file(path, header_text, records):
    """The records of process_regions through hipstr::VCFWriter (vcf_writer.h: POS-ordered heap; BGZF when the path ends in
    .gz, as the reference always writes) -- VCFWriter::open / write_header / add_vcf_record / close."""
    lib = capi.load()
    w = lib.hipstr_vcf_writer_open(path.encode())
    if not w:
        raise IOError("cannot create " + path)
    try:
        if lib.hipstr_vcf_writer_header(w, header_text.encode()) != 0:
            raise IOError("writing the header of %s failed" % path)
        for chrom, pos, text in records:
            if lib.hipstr_vcf_writer_add_record(w, chrom.encode(), pos, text.encode()) != 0:
                raise IOError("writing a record of %s failed" % path)
    finally:
        lib.hipstr_vcf_writer_close(w)
