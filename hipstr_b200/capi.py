"""ctypes view of the C-ABI in include/hipstr_b200.h (libhipstr_b200.so) and of the
synthetic generator (libhipstr_synth.so).

Python is only the test / bench driver here: the product is the shared library.  No
fallback exists -- `load()` raises if the library was not built (run
`__graft_entry__.build()` or `make -C hipstr_b200/csrc`), and `Context()` raises if no
CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhipstr_b200.so")
SYNTH_PATH = os.path.join(_HERE, "libhipstr_synth.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)


class AlignBatch(C.Structure):
    """hipstr_align_batch_t"""
    _fields_ = [
        ("n_loci", C.c_int32), ("n_blocks", C.c_int32), ("n_options", C.c_int32), ("n_pools", C.c_int32),
        ("n_haps", C.c_int64),
        ("locus_block_off", c_i32p), ("locus_pool_off", c_i32p), ("locus_hap_off", c_i64p), ("locus_out_off", c_i64p),
        ("block_period", c_i32p), ("block_opt_off", c_i32p), ("block_stutter", c_f64p),
        ("opt_seq_off", c_i32p), ("opt_seq", C.c_char_p),
        ("pool_seq_off", c_i32p), ("pool_bases", C.c_char_p), ("pool_quals", C.c_char_p), ("pool_seed", c_i32p),
        ("realign_pool", c_u8p), ("realign_hap", c_u8p),
    ]


class SynthCfg(C.Structure):
    _fields_ = [
        ("n_loci", C.c_int32), ("n_samples", C.c_int32), ("reads_per_sample", C.c_int32), ("n_alleles", C.c_int32),
        ("read_len", C.c_int32), ("trim", C.c_int32), ("period", C.c_int32), ("ref_copies", C.c_int32),
        ("seed", C.c_uint64), ("stutter_rate", C.c_double), ("sub_rate", C.c_double), ("mate_rate", C.c_double),
        ("flank_snp_freq", C.c_double), ("haploid", C.c_int32),
    ]


class SynthView(C.Structure):
    _fields_ = [
        ("batch", AlignBatch), ("n_reads", C.c_int64),
        ("locus_read_off", c_i32p), ("locus_sample_off", c_i32p), ("pool_index", c_i32p), ("sample_label", c_i32p),
        ("second_mate", c_u8p), ("read_weight", c_i32p), ("log_p1", c_f64p), ("log_p2", c_f64p),
        ("n_haps", c_i32p), ("haploid", c_u8p), ("true_gt", c_i32p), ("read_bp_diff", c_i32p),
        ("read_ll_size", C.c_int64), ("post_size", C.c_int64),
        ("read_seq_off", c_i32p), ("read_bases", C.c_void_p), ("read_quals", C.c_void_p), ("read_start", c_i32p),
        ("read_cigar_off", c_i32p), ("read_cigar_type", C.c_void_p), ("read_cigar_len", c_i32p),
        ("read_name_id", c_i32p), ("block_start", c_i32p), ("block_end", c_i32p), ("chrom_len", C.c_int32),
        ("chrom_seqs", C.c_void_p), ("region_start", C.c_int32), ("region_stop", C.c_int32), ("read_stop", c_i32p), ("read_rev_strand", c_u8p),
    ]


class ReadsBatch(C.Structure):
    """hipstr_reads_batch_t"""
    _fields_ = [
        ("locus_read_off", c_i32p), ("locus_sample_off", c_i32p), ("pool_index", c_i32p), ("sample_label", c_i32p),
        ("second_mate", c_u8p), ("read_weight", c_i32p), ("log_p1", c_f64p), ("log_p2", c_f64p), ("haploid", c_u8p),
        ("copy_read", c_u8p),
    ]


class LocusReadsStruct(C.Structure):
    """hipstr_locus_reads_t"""
    _fields_ = [("locus_read_off", c_i32p), ("locus_sample_off", c_i32p), ("read_seq_off", c_i32p), ("bases", C.c_void_p),
                ("quals", C.c_void_p), ("read_start", c_i32p), ("cigar_off", c_i32p), ("cigar_type", C.c_void_p),
                ("cigar_len", c_i32p), ("sample_label", c_i32p), ("name_id", c_i32p), ("log_p1", c_f64p), ("log_p2", c_f64p),
                ("haploid", c_u8p), ("rev_strand", c_u8p), ("read_stop", c_i32p), ("use_for_haps", c_u8p)]


class SnpPhasingStruct(C.Structure):
    """hipstr_snp_phasing_t"""
    _fields_ = [("n_entries", C.c_int32), ("entry_aln_off", c_i32p), ("entry_snp_set", c_i32p), ("n_alns", C.c_int32),
                ("aln_pos", c_i32p), ("aln_end", c_i32p), ("aln_seq_off", c_i32p), ("bases", C.c_void_p), ("quals", C.c_void_p),
                ("aln_cigar_off", c_i32p), ("cigar_type", C.c_void_p), ("cigar_len", c_i32p), ("n_sets", C.c_int32),
                ("set_off", c_i32p), ("snp_pos", C.c_void_p), ("snp_base1", C.c_void_p), ("snp_base2", C.c_void_p)]


class SnpPhasing:
    """Flat arrays of a hipstr_snp_phasing_t built from Python lists.

    entries: [(snp_set or -1, [alignment, ...])] with alignment = (pos, end, bases, quals, [(cigar char, length), ...]);
    sets: [[(pos, base one, base two), ...]] sorted by position."""

    def __init__(self, entries, sets):
        alns = [a for _, al in entries for a in al]
        self.entry_aln_off = np.zeros(len(entries) + 1, np.int32)
        self.entry_aln_off[1:] = np.cumsum([len(al) for _, al in entries])
        self.entry_snp_set = np.array([s for s, _ in entries], np.int32).reshape(-1)
        self.aln_pos = np.array([a[0] for a in alns], np.int32).reshape(-1)
        self.aln_end = np.array([a[1] for a in alns], np.int32).reshape(-1)
        self.aln_seq_off = np.zeros(len(alns) + 1, np.int32)
        self.aln_seq_off[1:] = np.cumsum([len(a[2]) for a in alns])
        self.bases = np.frombuffer(b"".join(_as_bytes(a[2]) for a in alns) + b"\0", np.uint8).copy()
        self.quals = np.frombuffer(b"".join(_as_bytes(a[3]) for a in alns) + b"\0", np.uint8).copy()
        self.aln_cigar_off = np.zeros(len(alns) + 1, np.int32)
        self.aln_cigar_off[1:] = np.cumsum([len(a[4]) for a in alns])
        self.cigar_type = np.frombuffer("".join(t for a in alns for t, _ in a[4]).encode() + b"\0", np.uint8).copy()
        self.cigar_len = np.array([n for a in alns for _, n in a[4]] + [0], np.int32)
        self.set_off = np.zeros(len(sets) + 1, np.int32)
        self.set_off[1:] = np.cumsum([len(x) for x in sets])
        snps = [x for st in sets for x in st]
        self.snp_pos = np.array([x[0] for x in snps] + [0], np.uint32)
        self.snp_base1 = np.frombuffer("".join(x[1] for x in snps).encode() + b"\0", np.uint8).copy()
        self.snp_base2 = np.frombuffer("".join(x[2] for x in snps).encode() + b"\0", np.uint8).copy()
        self.n_entries, self.n_alns, self.n_sets = len(entries), len(alns), len(sets)
        self.struct = SnpPhasingStruct(
            self.n_entries, ptr(self.entry_aln_off, c_i32p), ptr(self.entry_snp_set, c_i32p), self.n_alns,
            ptr(self.aln_pos, c_i32p), ptr(self.aln_end, c_i32p), ptr(self.aln_seq_off, c_i32p), self.bases.ctypes.data,
            self.quals.ctypes.data, ptr(self.aln_cigar_off, c_i32p), self.cigar_type.ctypes.data, ptr(self.cigar_len, c_i32p),
            self.n_sets, ptr(self.set_off, c_i32p), self.snp_pos.ctypes.data, self.snp_base1.ctypes.data,
            self.snp_base2.ctypes.data)

    @classmethod
    def from_arrays(cls, entry_aln_off, entry_snp_set, aln_pos, aln_end, aln_seq_off, bases, quals, aln_cigar_off, cigar_type, cigar_len,
                    set_off, snp_pos, snp_base1, snp_base2):
        """The same batch from flat numpy arrays (bases / quals / cigar_type / snp_base* as uint8 arrays or bytes)."""
        b = cls.__new__(cls)
        u8 = lambda x: np.frombuffer(bytes(x) + b"\0", np.uint8).copy() if isinstance(x, (bytes, bytearray)) else np.ascontiguousarray(
            np.concatenate([np.asarray(x, np.uint8), np.zeros(1, np.uint8)]))
        i32 = lambda x: np.ascontiguousarray(x, np.int32)
        b.entry_aln_off, b.entry_snp_set = i32(entry_aln_off), i32(entry_snp_set)
        b.aln_pos, b.aln_end, b.aln_seq_off = i32(aln_pos), i32(aln_end), i32(aln_seq_off)
        b.bases, b.quals = u8(bases), u8(quals)
        b.aln_cigar_off, b.cigar_type = i32(aln_cigar_off), u8(cigar_type)
        b.cigar_len = np.ascontiguousarray(np.concatenate([np.asarray(cigar_len, np.int32), np.zeros(1, np.int32)]))
        b.set_off = i32(set_off)
        b.snp_pos = np.ascontiguousarray(np.concatenate([np.asarray(snp_pos, np.uint32), np.zeros(1, np.uint32)]))
        b.snp_base1, b.snp_base2 = u8(snp_base1), u8(snp_base2)
        b.n_entries, b.n_alns, b.n_sets = len(b.entry_snp_set), len(b.aln_pos), len(b.set_off) - 1
        b.struct = SnpPhasingStruct(
            b.n_entries, ptr(b.entry_aln_off, c_i32p), ptr(b.entry_snp_set, c_i32p), b.n_alns, ptr(b.aln_pos, c_i32p), ptr(b.aln_end, c_i32p),
            ptr(b.aln_seq_off, c_i32p), b.bases.ctypes.data, b.quals.ctypes.data, ptr(b.aln_cigar_off, c_i32p), b.cigar_type.ctypes.data,
            ptr(b.cigar_len, c_i32p), b.n_sets, ptr(b.set_off, c_i32p), b.snp_pos.ctypes.data, b.snp_base1.ctypes.data,
            b.snp_base2.ctypes.data)
        return b

    def run(self, fn, *handle):
        """Calls a function with the product's signature (ctx?, batch, log_p1, log_p2, counts) -> (status, p1, p2, counts)."""
        p1, p2 = np.zeros(max(self.n_entries, 1)), np.zeros(max(self.n_entries, 1))
        counts = np.zeros((max(self.n_entries, 1), 4), np.int32)
        st = fn(*handle, C.byref(self.struct), ptr(p1, c_f64p), ptr(p2, c_f64p), ptr(counts, c_i32p))
        return st, p1[:self.n_entries], p2[:self.n_entries], counts[:self.n_entries]


def _as_bytes(x):
    return x if isinstance(x, (bytes, bytearray)) else x.encode("latin-1")


class FilterOptions(C.Structure):
    """hipstr_filter_options_t"""
    _fields_ = [("max_mate_dist", C.c_int32), ("min_bp_before_indel", C.c_int32), ("min_flank", C.c_int32),
                ("min_read_end_match", C.c_int32), ("maximal_end_match_window", C.c_int32), ("require_paired_reads", C.c_int32),
                ("min_sum_qual_log_prob", C.c_double), ("max_total_reads", C.c_int32), ("base_qual_trim", C.c_int32),
                ("remove_pcr_dups", C.c_int32), ("trim_adapters", C.c_int32)]


class PipelineOptions(C.Structure):
    """hipstr_pipeline_options_t"""
    _fields_ = [("filter", FilterOptions), ("max_str_length", C.c_int32), ("min_total_reads", C.c_int32), ("max_total_haplotypes", C.c_int32),
                ("max_flank_haplotypes", C.c_int32), ("min_flank_freq", C.c_double), ("max_em_iter", C.c_int32), ("abs_ll_converge", C.c_double),
                ("frac_ll_converge", C.c_double), ("use_def_stutter_model", C.c_int32), ("def_stutter_model", C.c_double * 6),
                ("recalc_stutter_model", C.c_int32), ("skip_padding", C.c_int32), ("n_haploid_chroms", C.c_int32),
                ("haploid_chroms", C.POINTER(C.c_char_p)), ("host_threads", C.c_int32), ("bams_from_10x", C.c_int32),
                ("ref_vcf", C.c_void_p)]


class FilteredView(C.Structure):
    """hipstr_filtered_view_t"""
    _fields_ = [("n_samples", C.c_int32), ("sample_names", C.POINTER(C.c_char_p)), ("sample_entry_off", c_i32p),
                ("entry_passes", C.c_void_p), ("aln_flag", c_i32p), ("entry_name_off", c_i32p), ("entry_names", C.c_void_p),
                ("reads", SnpPhasingStruct)]


def _text(fn, h):
    """Calls a *_text(handle, cap, out) function, growing the buffer to the size it asks for."""
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = fn(h, cap, buf)
        if n >= 0:
            return buf.raw[:n].decode("latin-1")
        if -n <= cap:
            raise HipstrError(-2, "text call failed")
        cap = -n


class SnpVcf:
    """hipstr_snp_vcf_t: a phased SNP VCF, loaded once; region queries return per-sample SNP sets for K7."""

    def __init__(self, path):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.hipstr_snp_vcf_open(path.encode(), C.byref(h))
        if st != 0:
            raise HipstrError(st, "snp_vcf_open: " + self.lib.hipstr_snp_vcf_last_error().decode())
        self.h = h
        self.samples = self.lib.hipstr_snp_vcf_samples(h).decode().splitlines()

    def has_chromosome(self, chrom):
        return bool(self.lib.hipstr_snp_vcf_has_chromosome(self.h, chrom.encode()))

    def region_sets(self, chrom, start, end, skip_regions=(), skip_padding=15):
        """-> None if the chromosome is absent, else (set_off [n_samples+1], pos uint32, base1 bytes, base2 bytes) (copies)."""
        ss = np.array([r[0] for r in skip_regions], np.int32)
        se = np.array([r[1] for r in skip_regions], np.int32)
        found = C.c_int32()
        off, pos, b1, b2 = c_i32p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        st = self.lib.hipstr_snp_vcf_region_sets(self.h, chrom.encode(), start, end, len(ss), ptr(ss, c_i32p) if len(ss) else None,
                                                 ptr(se, c_i32p) if len(se) else None, skip_padding, C.byref(found), C.byref(off),
                                                 C.byref(pos), C.byref(b1), C.byref(b2))
        if st != 0:
            raise HipstrError(st, "snp_vcf_region_sets")
        if not found.value:
            return None
        n = len(self.samples)
        set_off = np.ctypeslib.as_array(off, shape=(n + 1,)).copy()
        total = int(set_off[-1])
        p = np.frombuffer(C.string_at(pos.value, 4 * total), np.uint32).copy() if total else np.zeros(0, np.uint32)
        return set_off, p, C.string_at(b1.value, total), C.string_at(b2.value, total)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_snp_vcf_close(self.h)
            self.h = None

    __del__ = close


def vcf_header(reference_path, full_command, contigs, samples, **options):
    """hipstr_vcf_header: contigs = [(name, length)] in FASTA order; options override hipstr_vcf_default_options."""
    lib = load()
    opt = VcfOptions()
    lib.hipstr_vcf_default_options(C.byref(opt))
    for k, v in options.items():
        setattr(opt, k, v)
    mk = lambda xs: (C.c_char_p * max(len(xs), 1))(*[x.encode() for x in xs])
    lens = np.array([c[1] for c in contigs] + [0], np.int64)
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = lib.hipstr_vcf_header(reference_path.encode(), full_command.encode(), len(contigs), mk([c[0] for c in contigs]), ptr(lens, c_i64p),
                                  len(samples), mk(list(samples)), C.byref(opt), cap, buf)
        if n > 0:
            return buf.raw[:n].decode()
        if n == 0:
            raise HipstrError(-2, "vcf_header")
        cap = -n


class StrVcf:
    """hipstr_str_vcf_t: a reference panel of STR genotypes (--ref-vcf)."""

    def __init__(self, path):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.hipstr_str_vcf_open(path.encode(), C.byref(h))
        if st != 0:
            raise HipstrError(st, "str_vcf_open: " + self.lib.hipstr_snp_vcf_last_error().decode())
        self.h = h

    def alleles(self, chrom, region_start, region_stop):
        """read_vcf_alleles -> (pos, [alleles]) or None"""
        pos, n, text = C.c_int32(), C.c_int32(), C.c_char_p()
        if self.lib.hipstr_str_vcf_alleles(self.h, chrom.encode(), region_start, region_stop, C.byref(pos), C.byref(n), C.byref(text)) != 1:
            return None
        return pos.value, text.value.decode().splitlines()

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_str_vcf_close(self.h)
            self.h = None

    __del__ = close


class BamReader:
    """hipstr_bam_reader_t: BAM files (+ .bai) read region by region, file after file."""

    def __init__(self, paths):
        self.lib, self.paths = load(), list(paths)
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        h = C.c_void_p()
        st = self.lib.hipstr_bam_reader_open(len(paths), arr, C.byref(h))
        if st != 0:
            raise HipstrError(st, "bam_reader_open: " + self.lib.hipstr_ingest_last_error().decode())
        self.h = h

    def read_groups(self):
        """[(path, id, sample or None, library or None)]"""
        rows = [line.split("\t") for line in _text(self.lib.hipstr_bam_reader_read_groups, self.h).splitlines()]
        return [(r[0], r[1], None if r[2] == "-" else r[2], None if r[3] == "-" else r[3]) for r in rows]

    def fetch(self, chrom, start, end):
        h = C.c_void_p()
        st = self.lib.hipstr_bam_reader_fetch(self.h, chrom.encode(), start, end, C.byref(h))
        if st != 0:
            raise HipstrError(st, "bam_reader_fetch: " + self.lib.hipstr_ingest_last_error().decode())
        return BamRecords(self.lib, h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_bam_reader_close(self.h)
            self.h = None

    __del__ = close


class BamRecords:
    """hipstr_bam_records_t"""

    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def __len__(self):
        return self.lib.hipstr_bam_records_count(self.h)

    def text(self):
        return _text(self.lib.hipstr_bam_records_text, self.h)

    def filter(self, chrom_seq, regions, rg_map, options=None, **overrides):
        """hipstr_filter_reads; regions = [(start, stop)], rg_map = {file path + read group id: (sample, library)}."""
        opt = options or FilterOptions()
        if options is None:
            self.lib.hipstr_filter_default_options(C.byref(opt))
        for k, v in overrides.items():
            setattr(opt, k, v)
        starts = np.array([r[0] for r in regions], np.int32)
        stops = np.array([r[1] for r in regions], np.int32)
        keys = list(rg_map)
        mk = lambda xs: (C.c_char_p * max(len(xs), 1))(*[x.encode() for x in xs])
        h = C.c_void_p()
        seq = chrom_seq if isinstance(chrom_seq, bytes) else chrom_seq.encode()
        st = self.lib.hipstr_filter_reads(self.h, seq, len(regions), ptr(starts, c_i32p), ptr(stops, c_i32p), C.byref(opt), len(keys),
                                          mk(keys), mk([rg_map[k][0] for k in keys]), mk([rg_map[k][1] for k in keys]), C.byref(h))
        if st != 0:
            raise HipstrError(st, "filter_reads: " + self.lib.hipstr_ingest_last_error().decode())
        return FilteredReads(self.lib, h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_bam_records_free(self.h)
            self.h = None

    __del__ = close


class FilteredReads:
    """hipstr_filtered_reads_t"""
    COUNTS = ("overlapping", "hard_clipped", "has_n", "low_quality", "no_unique_mapping", "no_mate", "pcr_duplicates", "too_many_reads",
              "passed")

    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def text(self):
        return _text(self.lib.hipstr_filtered_reads_text, self.h)

    def counts(self):
        c = np.zeros(9, np.int32)
        self.lib.hipstr_filtered_reads_counts(self.h, ptr(c, c_i32p))
        return dict(zip(self.COUNTS, map(int, c)))

    def view(self):
        """FilteredView (pointers owned by this object, valid until the next view() / close())."""
        v = FilteredView()
        st = self.lib.hipstr_filtered_reads_view(self.h, C.byref(v))
        if st != 0:
            raise HipstrError(st, "filtered_reads_view")
        return v

    def entry_names(self):
        """Read names of the STR reads, in entry order."""
        v = self.view()
        n = v.reads.n_entries
        off = np.ctypeslib.as_array(v.entry_name_off, shape=(n + 1,))
        blob = C.string_at(v.entry_names, int(off[-1])).decode("latin-1")
        return [blob[off[i]:off[i + 1]] for i in range(n)]

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_filtered_reads_free(self.h)
            self.h = None

    __del__ = close


class VcfLoci(C.Structure):
    """hipstr_vcf_loci_t"""
    _fields_ = [("chrom", C.POINTER(C.c_char_p)), ("name", C.POINTER(C.c_char_p)), ("region_start", c_i32p),
                ("region_stop", c_i32p), ("period", c_i32p), ("chrom_seq", C.POINTER(C.c_char_p)),
                ("locus_sample_names", C.POINTER(C.c_char_p)), ("n_out_samples", C.c_int32),
                ("out_sample_names", C.POINTER(C.c_char_p))]


class VcfOptions(C.Structure):
    """hipstr_vcf_options_t"""
    _fields_ = [("output_gls", C.c_int32), ("output_pls", C.c_int32), ("output_phased_gls", C.c_int32),
                ("output_allreads", C.c_int32), ("output_mallreads", C.c_int32), ("output_filters", C.c_int32),
                ("output_haplotype_data", C.c_int32), ("max_flank_indel_frac", C.c_double)]


class GenotypeOut(C.Structure):
    """hipstr_genotype_out_t (host or device pointers)"""
    _fields_ = [("read_ll", C.c_void_p), ("read_seed", C.c_void_p), ("post", C.c_void_p), ("sample_ll", C.c_void_p),
                ("best", C.c_void_p), ("total_ll", C.c_void_p)]


class EmBatch(C.Structure):
    """hipstr_em_batch_t"""
    _fields_ = [("n_loci", C.c_int32), ("locus_read_off", c_i32p), ("locus_sample_off", c_i32p), ("num_bps", c_i32p),
                ("sample_label", c_i32p), ("log_p1", c_f64p), ("log_p2", c_f64p), ("motif_len", c_i32p),
                ("ref_allele", c_i32p), ("haploid", c_u8p)]


def make_em_batch(locus_read_off, locus_sample_off, num_bps, sample_label, log_p1, log_p2, motif_len, ref_allele, haploid):
    arrs = [np.ascontiguousarray(a, dt) for a, dt in ((locus_read_off, np.int32), (locus_sample_off, np.int32), (num_bps, np.int32),
            (sample_label, np.int32), (log_p1, np.float64), (log_p2, np.float64), (motif_len, np.int32), (ref_allele, np.int32),
            (haploid, np.uint8))]
    b = EmBatch(len(arrs[6]), ptr(arrs[0], c_i32p), ptr(arrs[1], c_i32p), ptr(arrs[2], c_i32p), ptr(arrs[3], c_i32p),
                ptr(arrs[4], c_f64p), ptr(arrs[5], c_f64p), ptr(arrs[6], c_i32p), ptr(arrs[7], c_i32p), ptr(arrs[8], c_u8p))
    b._keep = arrs
    return b


class TraceOut(C.Structure):
    """hipstr_trace_out_t"""
    _fields_ = [("aln_stride", C.c_int32), ("hap_aln", C.c_void_p), ("seed_hap_pos", c_i32p), ("stutter_size", c_i32p),
                ("span_start", c_i32p), ("span_len", c_i32p), ("flank_ins", c_i32p), ("flank_del", c_i32p),
                ("n_indels", c_i32p), ("indels", c_i32p), ("n_snps", c_i32p), ("snps", c_i32p)]


MAX_BLOCKS, MAX_TRACE_INDELS, MAX_TRACE_SNPS, NO_STR_DATA = 8, 16, 32, -2147483648


def trace_batch(fn, batch, block_start, trace_pool, trace_hap, aln_stride=1024, ctx_handle=None, extra_args=()):
    """Calls a trace_batch entry point; returns a dict of numpy arrays (hap_aln as a list of str)."""
    bs = np.ascontiguousarray(block_start, np.int32)
    tp = np.ascontiguousarray(trace_pool, np.int32)
    th = np.ascontiguousarray(trace_hap, np.int32)
    n = len(tp)
    o = dict(hap_aln=np.zeros(n * aln_stride, np.uint8), seed_hap_pos=np.zeros(n, np.int32),
             stutter_size=np.zeros(n * MAX_BLOCKS, np.int32), span_start=np.zeros(n * MAX_BLOCKS, np.int32),
             span_len=np.zeros(n * MAX_BLOCKS, np.int32), flank_ins=np.zeros(n, np.int32), flank_del=np.zeros(n, np.int32),
             n_indels=np.zeros(n, np.int32), indels=np.zeros(n * MAX_TRACE_INDELS * 2, np.int32),
             n_snps=np.zeros(n, np.int32), snps=np.zeros(n * MAX_TRACE_SNPS * 2, np.int32))
    to = TraceOut(aln_stride, o["hap_aln"].ctypes.data, *[ptr(o[k], c_i32p) for k in
                  ("seed_hap_pos", "stutter_size", "span_start", "span_len", "flank_ins", "flank_del", "n_indels", "indels",
                   "n_snps", "snps")])
    args = [C.byref(batch), ptr(bs, c_i32p), C.c_int32(n), ptr(tp, c_i32p), ptr(th, c_i32p), C.byref(to)] + list(extra_args)
    st = fn(ctx_handle, *args) if ctx_handle is not None else fn(*args)
    raw = o["hap_aln"].reshape(n, aln_stride)
    o["hap_aln"] = [bytes(r[:int(np.argmax(r == 0))]).decode() for r in raw]
    for k in ("stutter_size", "span_start", "span_len"):
        o[k] = o[k].reshape(n, MAX_BLOCKS)
    o["indels"] = o["indels"].reshape(n, MAX_TRACE_INDELS, 2)
    o["snps"] = o["snps"].reshape(n, MAX_TRACE_SNPS, 2)
    return st, o


def trace_flank_lists(lib, batch, block_start, pool, hap, hap_aln, seed_hap_pos, stutter_size, cap=4096, name="hipstr_trace_flank_lists"):
    """The complete flank indel / SNP lists of one trace (hipstr_trace_flank_lists, or the reference harness's
    ref_trace_lists when name says so).  Returns (indels [n][2], snps [n][2])."""
    bs = np.ascontiguousarray(block_start, np.int32)
    ind, snp = np.zeros(2 * cap, np.int32), np.zeros(2 * cap, np.int32)
    ni, ns = C.c_int32(), C.c_int32()
    f = getattr(lib, name)
    f.restype = C.c_int32
    if name == "hipstr_trace_flank_lists":
        ss = np.ascontiguousarray(stutter_size, np.int32)
        f.argtypes = [C.POINTER(AlignBatch), c_i32p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, c_i32p, C.c_char_p, C.c_int32,
                      c_i32p, c_i32p, C.c_int32, c_i32p, c_i32p]
        st = f(C.byref(batch), ptr(bs, c_i32p), int(pool), int(hap), hap_aln.encode(), int(seed_hap_pos), ptr(ss, c_i32p), None, cap,
               C.byref(ni), ptr(ind, c_i32p), cap, C.byref(ns), ptr(snp, c_i32p))
    else:
        f.argtypes = [C.POINTER(AlignBatch), c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p, c_i32p]
        st = f(C.byref(batch), ptr(bs, c_i32p), int(pool), int(hap), cap, C.byref(ni), ptr(ind, c_i32p), cap, C.byref(ns), ptr(snp, c_i32p))
    if st != 0:
        raise HipstrError(st, name)
    if ni.value > cap or ns.value > cap:   # too few slots: the counts are the true ones, ask again with room for them
        return trace_flank_lists(lib, batch, block_start, pool, hap, hap_aln, seed_hap_pos, stutter_size, max(ni.value, ns.value), name)
    return ind[:2 * ni.value].reshape(-1, 2).copy(), snp[:2 * ns.value].reshape(-1, 2).copy()


EXTRACT_ARGTYPES = [C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_u8p, c_f64p, c_f64p, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p,
                    c_f64p, c_f64p, c_f64p, c_f64p, c_i32p]


def extract_genotypes(fn, locus_sample_off, n_haps, n_variants, hap_to_allele, haploid, post, sample_ll, ctx_handle=None):
    """Calls an extract_genotypes entry point (product / oracle / reference harness share the signature)."""
    lso, nh, nv = (np.ascontiguousarray(a, np.int32) for a in (locus_sample_off, n_haps, n_variants))
    h2a = np.ascontiguousarray(hap_to_allele, np.int32)
    hp = np.ascontiguousarray(haploid, np.uint8)
    S, Sl = int(lso[-1]), np.diff(lso)
    G = np.where(hp != 0, nv, nv * (nv + 1) // 2)
    PG = np.where(hp != 0, nv, nv * nv)
    out = dict(best_hap=np.zeros(2 * S, np.int32), best_gt=np.zeros(2 * S, np.int32), log_phased=np.zeros(S),
               log_unphased=np.zeros(S), hap_log_phased=np.zeros(S), hap_log_unphased=np.zeros(S),
               gl=np.zeros(int((Sl * G).sum())), phased_gl=np.zeros(int((Sl * PG).sum())), gl_diff=np.zeros(S),
               pl=np.zeros(int((Sl * G).sum()), np.int32))
    args = [len(nh), ptr(lso, c_i32p), ptr(nh, c_i32p), ptr(nv, c_i32p), ptr(h2a, c_i32p), ptr(hp, c_u8p),
            ptr(np.ascontiguousarray(post), c_f64p), ptr(np.ascontiguousarray(sample_ll), c_f64p),
            ptr(out["best_hap"], c_i32p), ptr(out["best_gt"], c_i32p), ptr(out["log_phased"], c_f64p),
            ptr(out["log_unphased"], c_f64p), ptr(out["hap_log_phased"], c_f64p), ptr(out["hap_log_unphased"], c_f64p),
            ptr(out["gl"], c_f64p), ptr(out["phased_gl"], c_f64p), ptr(out["gl_diff"], c_f64p), ptr(out["pl"], c_i32p)]
    st = fn(ctx_handle, *args) if ctx_handle is not None else fn(*args)
    return st, out


def stitch_trace(hap_start, hap_aln_to_ref, read_aln_to_hap, seed_hap_pos, seed_base, read_bases):
    """hipstr_stitch_trace -> (status, start, stop, cigar string like '3S50M2D10M', gapped alignment)."""
    lib = load()
    cap = 4096
    start, stop, n = C.c_int32(), C.c_int32(), C.c_int32()
    ctype = C.create_string_buffer(cap)
    clen = np.zeros(cap, np.int32)
    aln = C.create_string_buffer(cap)
    st = lib.hipstr_stitch_trace(hap_start, hap_aln_to_ref.encode(), read_aln_to_hap.encode(), seed_hap_pos, seed_base,
                                 read_bases.encode(), C.byref(start), C.byref(stop), cap, ctype, ptr(clen, c_i32p), C.byref(n),
                                 cap, aln)
    cigar = "".join("%d%s" % (clen[i], ctype.raw[i:i + 1].decode()) for i in range(n.value))
    return st, start.value, stop.value, cigar, aln.value.decode()


def em_train(fn, batch, max_iter=100, min_abs=0.01, min_frac=0.001, ctx_handle=None):
    """Calls an em_train entry point (product, oracle or reference harness share the signature after the context)."""
    L = batch.n_loci
    prm = np.zeros(6 * L)
    conv = np.zeros(L, np.uint8)
    it = np.zeros(L, np.int32)
    ll = np.zeros(L)
    args = [C.byref(batch), C.c_int32(max_iter), C.c_double(min_abs), C.c_double(min_frac), ptr(prm, c_f64p), ptr(conv, c_u8p),
            ptr(it, c_i32p), ptr(ll, c_f64p)]
    st = fn(ctx_handle, *args) if ctx_handle is not None else fn(*args)
    return st, prm.reshape(L, 6), conv, it, ll


STATUS = {0: "OK", 1: "NO_DEVICE", 2: "CUDA", 3: "BAD_ARG", 4: "UNSUPPORTED", 5: "INVALID_SEED", 6: "BAD_CIGAR"}


class HipstrError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__("hipstr status %d (%s) %s" % (status, STATUS.get(status, "?"), msg))


def ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def bind_align_abi(lib, prefix):
    """Declare the entry points that take the flat batch (shared by product, oracle and reference harness)."""
    B = C.POINTER(AlignBatch)
    f = getattr(lib, prefix + "calc_seeds")
    f.restype = C.c_int32
    f.argtypes = [C.c_int32, c_i32p, c_i32p, c_i32p, C.c_char_p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p,
                  c_i32p, c_i32p]
    if prefix != "hipstr_":
        f = getattr(lib, prefix + "align_batch")
        f.restype = C.c_int32
        f.argtypes = [B, c_f64p, c_i32p]
        f = getattr(lib, prefix + "align_loci")
        f.restype = C.c_int32
        f.argtypes = [B, C.c_int32, C.c_int32, c_f64p, c_i32p]
        f = getattr(lib, prefix + "posteriors")
        f.restype = C.c_int32
        f.argtypes = [C.c_int32, c_i32p, c_i32p, c_i32p, c_u8p, c_f64p, c_f64p, c_f64p, c_i32p, c_i32p, c_f64p,
                      c_f64p, c_i32p, c_f64p]
    return lib


_lib = None
_synth = None
NEXT_WINDOW_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p)


def load():
    """Load libhipstr_b200.so; fails loudly when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    B = C.POINTER(AlignBatch)
    vp = C.c_void_p
    lib.hipstr_create.restype = C.c_int32
    lib.hipstr_create.argtypes = [C.c_int32, C.POINTER(vp)]
    lib.hipstr_destroy.restype = None
    lib.hipstr_destroy.argtypes = [vp]
    lib.hipstr_last_error.restype = C.c_char_p
    lib.hipstr_last_error.argtypes = [vp]
    lib.hipstr_version.restype = C.c_char_p
    lib.hipstr_set_stream.restype = C.c_int32
    lib.hipstr_set_stream.argtypes = [vp, vp]
    bind_align_abi(lib, "hipstr_")
    lib.hipstr_pool_reads.restype = C.c_int32
    lib.hipstr_pool_reads.argtypes = [C.c_int32, c_i32p, C.c_char_p, C.c_char_p, c_i32p, c_i32p, c_i32p, c_i32p,
                                      C.c_char_p, C.c_char_p]
    lib.hipstr_align_batch_host.restype = C.c_int32
    lib.hipstr_align_batch_host.argtypes = [vp, B, c_f64p, c_i32p]
    lib.hipstr_upload_batch.restype = C.c_int32
    lib.hipstr_upload_batch.argtypes = [vp, B, C.POINTER(vp)]
    lib.hipstr_align_batch_dev.restype = C.c_int32
    lib.hipstr_align_batch_dev.argtypes = [vp, vp, vp, vp]
    lib.hipstr_free_batch.restype = None
    lib.hipstr_free_batch.argtypes = [vp, vp]
    lib.hipstr_batch_num_alignments.restype = C.c_int64
    lib.hipstr_batch_num_alignments.argtypes = [B]
    lib.hipstr_last_launch_count.restype = C.c_int32
    lib.hipstr_last_launch_count.argtypes = [vp]
    lib.hipstr_enable_timing.restype = C.c_int32
    lib.hipstr_enable_timing.argtypes = [vp, C.c_int]
    lib.hipstr_last_kernel_ms.restype = C.c_float
    lib.hipstr_last_kernel_ms.argtypes = [vp]
    lib.hipstr_scatter_pool_lls_host.restype = C.c_int32
    lib.hipstr_scatter_pool_lls_host.argtypes = [vp, C.c_int32, C.c_int32, c_f64p, c_i32p, c_i32p, c_u8p, c_u8p,
                                                 c_u8p, c_f64p, c_i32p]
    lib.hipstr_posteriors_host.restype = C.c_int32
    lib.hipstr_posteriors_host.argtypes = [vp, C.c_int32, c_i32p, c_i32p, c_i32p, c_u8p, c_f64p, c_f64p, c_f64p,
                                           c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_f64p]
    RB, GO = C.POINTER(ReadsBatch), C.POINTER(GenotypeOut)
    lib.hipstr_genotype_batch_host.restype = C.c_int32
    lib.hipstr_genotype_batch_host.argtypes = [vp, B, RB, GO]
    lib.hipstr_upload_genotype_batch.restype = C.c_int32
    lib.hipstr_upload_genotype_batch.argtypes = [vp, B, RB, C.POINTER(vp)]
    lib.hipstr_genotype_batch_dev.restype = C.c_int32
    lib.hipstr_genotype_batch_dev.argtypes = [vp, vp, GO]
    lib.hipstr_free_genotype_batch.restype = None
    lib.hipstr_free_genotype_batch.argtypes = [vp, vp]
    lib.hipstr_extract_genotypes_host.restype = C.c_int32
    lib.hipstr_extract_genotypes_host.argtypes = [vp] + EXTRACT_ARGTYPES
    lib.hipstr_trace_batch_host.restype = C.c_int32
    lib.hipstr_trace_batch_host.argtypes = [vp, B, c_i32p, C.c_int32, c_i32p, c_i32p, C.POINTER(TraceOut)]
    lib.hipstr_vcf_writer_open.restype = vp
    lib.hipstr_vcf_writer_open.argtypes = [C.c_char_p]
    lib.hipstr_vcf_writer_header.restype = C.c_int32
    lib.hipstr_vcf_writer_header.argtypes = [vp, C.c_char_p]
    lib.hipstr_vcf_writer_add_record.restype = C.c_int32
    lib.hipstr_vcf_writer_add_record.argtypes = [vp, C.c_char_p, C.c_int32, C.c_char_p]
    lib.hipstr_vcf_writer_close.restype = None
    lib.hipstr_vcf_writer_close.argtypes = [vp]
    lib.hipstr_vcf_writer_finish.restype = C.c_int32
    lib.hipstr_vcf_writer_finish.argtypes = [vp]
    lib.hipstr_hap_aln_index.restype = C.c_int32
    lib.hipstr_hap_aln_index.argtypes = [C.c_char_p, C.c_int32, c_i32p]
    lib.hipstr_trace_span.restype = C.c_int32
    lib.hipstr_trace_span.argtypes = [C.c_int32, C.c_char_p, C.c_int32, c_i32p, C.c_char_p, C.c_int32, C.c_int32, c_i32p, c_i32p]
    lib.hipstr_stitch_trace.restype = C.c_int32
    lib.hipstr_stitch_trace.argtypes = [C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, c_i32p, c_i32p,
                                        C.c_int32, C.c_char_p, c_i32p, c_i32p, C.c_int32, C.c_char_p]
    lib.hipstr_em_train_host.restype = C.c_int32
    lib.hipstr_em_train_host.argtypes = [vp, C.POINTER(EmBatch), C.c_int32, C.c_double, C.c_double, c_f64p, c_u8p, c_i32p,
                                         c_f64p]
    lib.hipstr_trace_seconds.restype = None
    lib.hipstr_trace_seconds.argtypes = [vp, c_f64p]
    lib.hipstr_last_traffic.restype = None
    lib.hipstr_last_traffic.argtypes = [vp, c_i64p, c_i64p, c_i32p]
    cpp = C.POINTER(C.c_char_p)
    lib.hipstr_ingest_last_error.restype = C.c_char_p
    lib.hipstr_bam_reader_open.restype = C.c_int32
    lib.hipstr_bam_reader_open.argtypes = [C.c_int32, cpp, C.POINTER(vp)]
    lib.hipstr_bam_reader_close.restype = None
    lib.hipstr_bam_reader_close.argtypes = [vp]
    lib.hipstr_bam_reader_read_groups.restype = C.c_int64
    lib.hipstr_bam_reader_read_groups.argtypes = [vp, C.c_int64, C.c_char_p]
    lib.hipstr_bam_reader_fetch.restype = C.c_int32
    lib.hipstr_bam_reader_fetch.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.hipstr_bam_records_count.restype = C.c_int32
    lib.hipstr_bam_records_count.argtypes = [vp]
    lib.hipstr_bam_records_text.restype = C.c_int64
    lib.hipstr_bam_records_text.argtypes = [vp, C.c_int64, C.c_char_p]
    lib.hipstr_bam_records_free.restype = None
    lib.hipstr_bam_records_free.argtypes = [vp]
    lib.hipstr_filter_default_options.restype = None
    lib.hipstr_filter_default_options.argtypes = [C.POINTER(FilterOptions)]
    lib.hipstr_filter_reads.restype = C.c_int32
    lib.hipstr_filter_reads.argtypes = [vp, C.c_char_p, C.c_int32, c_i32p, c_i32p, C.POINTER(FilterOptions), C.c_int32, cpp, cpp, cpp,
                                        C.POINTER(vp)]
    lib.hipstr_filtered_reads_counts.restype = None
    lib.hipstr_filtered_reads_counts.argtypes = [vp, c_i32p]
    lib.hipstr_filtered_reads_text.restype = C.c_int64
    lib.hipstr_filtered_reads_text.argtypes = [vp, C.c_int64, C.c_char_p]
    lib.hipstr_filtered_reads_view.restype = C.c_int32
    lib.hipstr_filtered_reads_view.argtypes = [vp, C.POINTER(FilteredView)]
    lib.hipstr_filtered_reads_free.restype = None
    lib.hipstr_filtered_reads_free.argtypes = [vp]
    lib.hipstr_trim_one.restype = C.c_int32
    lib.hipstr_trim_one.argtypes = [C.c_int32] * 6 + [C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, c_i32p, C.c_char_p,
                                    C.c_char_p, c_i32p, C.c_char_p, c_i32p]
    lib.hipstr_alignment_filters.restype = C.c_int32
    lib.hipstr_alignment_filters.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, C.c_char_p,
                                             C.c_int32, c_i32p, c_f64p]
    lib.hipstr_vcf_header.restype = C.c_int64
    lib.hipstr_vcf_header.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, cpp, c_i64p, C.c_int32, cpp, C.POINTER(VcfOptions), C.c_int64,
                                      C.c_char_p]
    lib.hipstr_genotyper_create_with_ref_alleles.restype = C.c_int32
    lib.hipstr_genotyper_create_with_ref_alleles.argtypes = [vp, C.c_int32, c_i32p, c_i32p, c_i32p, cpp, c_f64p, C.POINTER(LocusReadsStruct),
                                                             c_i32p, c_i32p, cpp, C.POINTER(vp)]
    lib.hipstr_str_vcf_open.restype = C.c_int32
    lib.hipstr_str_vcf_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.hipstr_str_vcf_close.restype = None
    lib.hipstr_str_vcf_close.argtypes = [vp]
    lib.hipstr_str_vcf_alleles.restype = C.c_int32
    lib.hipstr_str_vcf_alleles.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, c_i32p, c_i32p, C.POINTER(C.c_char_p)]
    lib.hipstr_pipeline_default_options.restype = None
    lib.hipstr_pipeline_default_options.argtypes = [C.POINTER(PipelineOptions)]
    lib.hipstr_process_regions_last_error.restype = C.c_char_p
    lib.hipstr_process_regions.restype = C.c_int32
    lib.hipstr_process_regions.argtypes = [vp, C.c_int32, cpp, vp, C.c_int32, cpp, cpp, C.c_int32, cpp, c_i32p, c_i32p, c_i32p, cpp,
                                           C.POINTER(PipelineOptions), C.POINTER(VcfOptions), C.POINTER(vp)]
    lib.hipstr_region_results_count.restype = C.c_int32
    lib.hipstr_region_results_count.argtypes = [vp]
    lib.hipstr_region_results_status.restype = C.c_int32
    lib.hipstr_region_results_status.argtypes = [vp, C.c_int32, c_i32p, c_i32p]
    lib.hipstr_region_results_record.restype = C.c_char_p
    lib.hipstr_region_results_record.argtypes = [vp, C.c_int32]
    lib.hipstr_region_results_samples.restype = C.c_char_p
    lib.hipstr_region_results_samples.argtypes = [vp]
    lib.hipstr_region_results_timing.restype = None
    lib.hipstr_region_results_timing.argtypes = [vp, c_f64p, c_i64p]
    lib.hipstr_region_results_genotyper_timing.restype = None
    lib.hipstr_region_results_genotyper_timing.argtypes = [vp, c_f64p, c_i64p]
    lib.hipstr_region_results_free.restype = None
    lib.hipstr_region_results_free.argtypes = [vp]
    lib.hipstr_snp_vcf_last_error.restype = C.c_char_p
    lib.hipstr_snp_vcf_open.restype = C.c_int32
    lib.hipstr_snp_vcf_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.hipstr_snp_vcf_close.restype = None
    lib.hipstr_snp_vcf_close.argtypes = [vp]
    lib.hipstr_snp_vcf_num_samples.restype = C.c_int32
    lib.hipstr_snp_vcf_num_samples.argtypes = [vp]
    lib.hipstr_snp_vcf_samples.restype = C.c_char_p
    lib.hipstr_snp_vcf_samples.argtypes = [vp]
    lib.hipstr_snp_vcf_has_chromosome.restype = C.c_int32
    lib.hipstr_snp_vcf_has_chromosome.argtypes = [vp, C.c_char_p]
    lib.hipstr_snp_vcf_region_sets.restype = C.c_int32
    lib.hipstr_snp_vcf_region_sets.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p,
                                               C.POINTER(c_i32p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.hipstr_snp_phasing_batch_host.restype = C.c_int32
    lib.hipstr_snp_phasing_batch_host.argtypes = [vp, C.POINTER(SnpPhasingStruct), c_f64p, c_f64p, c_i32p]
    lib.hipstr_nw_align_batch_host.restype = C.c_int32
    lib.hipstr_nw_align_batch_host.argtypes = [vp, C.c_int32, c_i32p, C.c_char_p, c_i32p, C.c_char_p, C.c_int32, C.c_int32,
                                               C.c_void_p, c_i32p, C.POINTER(C.c_float)]
    lib.hipstr_left_align_reads_host.restype = C.c_int32
    lib.hipstr_left_align_reads_host.argtypes = [vp, C.c_int32, C.POINTER(LocusReadsStruct), C.POINTER(C.c_char_p), c_i32p, c_i32p,
                                                 C.POINTER(vp)]
    lib.hipstr_left_aligned_reads.restype = C.POINTER(LocusReadsStruct)
    lib.hipstr_left_aligned_reads.argtypes = [vp]
    lib.hipstr_left_aligned_source.restype = c_i32p
    lib.hipstr_left_aligned_source.argtypes = [vp, c_i64p]
    lib.hipstr_left_aligned_counts.restype = None
    lib.hipstr_left_aligned_counts.argtypes = [vp, c_i64p, c_i64p]
    lib.hipstr_left_aligned_free.restype = None
    lib.hipstr_left_aligned_free.argtypes = [vp]
    lib.hipstr_left_align_one.restype = C.c_int32
    lib.hipstr_left_align_one.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, c_i32p, C.c_char_p,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_char_p, c_i32p, c_i32p, C.c_char_p, C.c_char_p,
                                          c_i32p, C.c_char_p, c_i32p]
    lib.hipstr_hap_aln_to_ref.restype = C.c_int32
    lib.hipstr_hap_aln_to_ref.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p]
    lib.hipstr_genotyper_create.restype = C.c_int32
    lib.hipstr_genotyper_create.argtypes = [vp, B, c_i32p, c_i32p, C.POINTER(LocusReadsStruct), C.POINTER(vp)]
    lib.hipstr_genotyper_create_from_reads.restype = C.c_int32
    lib.hipstr_genotyper_create_from_reads.argtypes = [vp, C.c_int32, c_i32p, c_i32p, c_i32p, C.POINTER(C.c_char_p), c_f64p,
                                                       C.POINTER(LocusReadsStruct), C.POINTER(vp)]
    lib.hipstr_genotyper_destroy.restype = None
    lib.hipstr_genotyper_destroy.argtypes = [vp]
    lib.hipstr_genotyper_last_error.restype = C.c_char_p
    lib.hipstr_genotyper_last_error.argtypes = [vp]
    lib.hipstr_genotyper_genotype.restype = C.c_int32
    lib.hipstr_genotyper_genotype.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, c_u8p]
    lib.hipstr_genotyper_recompute_stutter_models.restype = C.c_int32
    lib.hipstr_genotyper_recompute_stutter_models.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double,
                                                              C.c_double, c_u8p]
    lib.hipstr_genotyper_stats.restype = C.c_int32
    lib.hipstr_genotyper_stats.argtypes = [vp, c_i64p, c_i64p, c_i32p]
    lib.hipstr_genotyper_timing.restype = C.c_int32
    lib.hipstr_genotyper_timing.argtypes = [vp, c_f64p]
    lib.hipstr_genotyper_phase_timing.restype = C.c_int32
    lib.hipstr_genotyper_phase_timing.argtypes = [vp, c_f64p]
    lib.hipstr_genotyper_locus_info.restype = C.c_int32
    lib.hipstr_genotyper_locus_info.argtypes = [vp, C.c_int32, c_i32p]
    lib.hipstr_genotyper_locus_blocks.restype = C.c_int32
    lib.hipstr_genotyper_locus_blocks.argtypes = [vp, C.c_int32, c_i32p, c_i32p, C.c_void_p]
    lib.hipstr_genotyper_locus_results.restype = C.c_int32
    lib.hipstr_genotyper_locus_results.argtypes = [vp, C.c_int32, c_f64p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_u8p]
    lib.hipstr_vcf_default_options.restype = None
    lib.hipstr_vcf_default_options.argtypes = [C.POINTER(VcfOptions)]
    lib.hipstr_genotyper_write_vcf.restype = C.c_int32
    lib.hipstr_genotyper_write_vcf.argtypes = [vp, C.POINTER(VcfLoci), C.POINTER(VcfOptions)]
    lib.hipstr_genotyper_locus_record.restype = C.c_int32
    lib.hipstr_genotyper_locus_record.argtypes = [vp, C.c_int32, c_i32p, C.c_void_p, C.c_int32]
    lib.hipstr_genotyper_emit_records.restype = C.c_int32
    lib.hipstr_genotyper_emit_records.argtypes = [vp, C.POINTER(VcfLoci), vp]
    lib.hipstr_genotyper_locus_log.restype = C.c_int32
    lib.hipstr_genotyper_locus_log.argtypes = [vp, C.c_int32, C.c_void_p, C.c_int32]
    lib.hipstr_collect_timing.restype = C.c_int32
    lib.hipstr_collect_timing.argtypes = [vp, c_f64p, c_f64p, c_i32p]
    lib.hipstr_multi_create.restype = C.c_int32
    lib.hipstr_multi_create.argtypes = [C.c_int32, c_i32p, C.c_int32, C.POINTER(vp)]
    lib.hipstr_multi_destroy.restype = None
    lib.hipstr_multi_destroy.argtypes = [vp]
    lib.hipstr_multi_last_error.restype = C.c_char_p
    lib.hipstr_multi_last_error.argtypes = [vp]
    lib.hipstr_multi_num_workers.restype = C.c_int32
    lib.hipstr_multi_num_workers.argtypes = [vp]
    lib.hipstr_multi_num_windows.restype = C.c_int32
    lib.hipstr_multi_num_windows.argtypes = [C.c_int32, C.c_int32]
    lib.hipstr_multi_window_order.restype = C.c_int32
    lib.hipstr_multi_window_order.argtypes = [C.c_int32, C.c_int32, c_i32p, c_i32p]
    lib.hipstr_multi_genotype.restype = C.c_int32
    lib.hipstr_multi_genotype.argtypes = [vp, C.c_int32, c_i32p, c_i32p, c_i32p, C.POINTER(C.c_char_p), c_f64p, C.POINTER(LocusReadsStruct),
                                          C.POINTER(VcfLoci), C.POINTER(VcfOptions), C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                          NEXT_WINDOW_FN, vp, c_u8p]
    lib.hipstr_multi_locus_record.restype = C.c_int32
    lib.hipstr_multi_locus_record.argtypes = [vp, C.c_int32, c_i32p, vp, C.c_int32]
    lib.hipstr_multi_stats.restype = C.c_int32
    lib.hipstr_multi_stats.argtypes = [vp, c_i64p, c_i64p, c_f64p, c_i32p, c_f64p]
    lib.hipstr_multi_traffic.restype = C.c_int32
    lib.hipstr_multi_traffic.argtypes = [vp, c_i64p, c_i64p, c_i64p]
    _lib = lib
    return lib


def load_synth():
    global _synth
    if _synth is not None:
        return _synth
    if not os.path.exists(SYNTH_PATH):
        raise ImportError("%s missing: build it with `make -C hipstr_b200/csrc`" % SYNTH_PATH)
    lib = C.CDLL(SYNTH_PATH)
    lib.hipstr_synth_create.restype = C.c_void_p
    lib.hipstr_synth_create.argtypes = [C.POINTER(SynthCfg)]
    lib.hipstr_synth_view.restype = C.POINTER(SynthView)
    lib.hipstr_synth_view.argtypes = [C.c_void_p]
    lib.hipstr_synth_destroy.restype = None
    lib.hipstr_synth_destroy.argtypes = [C.c_void_p]
    _synth = lib
    return lib


def _np(p, n, dtype):
    if n == 0 or not p:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).view(dtype)


class Synth:
    """A batch of synthetic loci owned by libhipstr_synth.so."""

    def __init__(self, n_loci, n_samples, reads_per_sample, n_alleles, read_len, seed=1, trim=1, period=4,
                 ref_copies=12, stutter_rate=0.05, sub_rate=1.0 / 200, mate_rate=0.0, flank_snp_freq=0.0, haploid=0):
        lib = load_synth()
        self.cfg = SynthCfg(n_loci, n_samples, reads_per_sample, n_alleles, read_len, trim, period, ref_copies, seed,
                            stutter_rate, sub_rate, mate_rate, flank_snp_freq, haploid)
        self._h = lib.hipstr_synth_create(C.byref(self.cfg))
        if not self._h:
            raise HipstrError(3, "hipstr_synth_create: the reads of %d loci do not fit the 32-bit offsets of one batch" % n_loci)
        self.view = lib.hipstr_synth_view(self._h).contents
        v, b = self.view, self.view.batch
        self.batch = b
        self.n_loci = b.n_loci
        self.n_pools = b.n_pools
        self.n_reads = v.n_reads
        L = b.n_loci
        self.locus_pool_off = _np(b.locus_pool_off, L + 1, np.int32)
        self.locus_out_off = _np(b.locus_out_off, L + 1, np.int64)
        self.locus_hap_off = _np(b.locus_hap_off, L + 1, np.int64)
        self.pool_seq_off = _np(b.pool_seq_off, b.n_pools + 1, np.int32)
        self.pool_seed = _np(b.pool_seed, b.n_pools, np.int32)
        self.locus_read_off = _np(v.locus_read_off, L + 1, np.int32)
        self.locus_sample_off = _np(v.locus_sample_off, L + 1, np.int32)
        self.pool_index = _np(v.pool_index, v.n_reads, np.int32)
        self.sample_label = _np(v.sample_label, v.n_reads, np.int32)
        self.second_mate = _np(v.second_mate, v.n_reads, np.uint8)
        self.read_weight = _np(v.read_weight, v.n_reads, np.int32)
        self.log_p1 = _np(v.log_p1, v.n_reads, np.float64)
        self.log_p2 = _np(v.log_p2, v.n_reads, np.float64)
        self.n_haps = _np(v.n_haps, L, np.int32)
        self.haploid = _np(v.haploid, L, np.uint8)
        self.n_out = int(self.locus_out_off[-1])
        self.read_ll_size = v.read_ll_size
        self.post_size = v.post_size

    def reads_batch(self, copy_read=None):
        """hipstr_reads_batch_t view of the read-level arrays."""
        v = self.view
        rb = ReadsBatch(v.locus_read_off, v.locus_sample_off, v.pool_index, v.sample_label, v.second_mate,
                        v.read_weight, v.log_p1, v.log_p2, v.haploid, ptr(copy_read, c_u8p))
        rb._keep = copy_read
        return rb

    def close(self):
        if self._h:
            load_synth().hipstr_synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hap_aln_to_ref(ref_hap, alt_hap, first_block_start, repeat_block_start):
    """hipstr_hap_aln_to_ref -> 'M'/'I'/'D' string (Haplotype::aln_haps_to_ref for one haplotype)."""
    lib = load()
    cap = len(ref_hap) + len(alt_hap) + 8
    buf = C.create_string_buffer(cap)
    st = lib.hipstr_hap_aln_to_ref(ref_hap.encode(), alt_hap.encode(), first_block_start, repeat_block_start, cap, buf)
    if st != 0:
        raise HipstrError(st, "hap_aln_to_ref")
    return buf.value.decode()


def blocks_batch(loci_blocks, stutter=(0.95, 0.05, 0.05, 0.95, 0.01, 0.01)):
    """hipstr_align_batch_t carrying only haplotype blocks: loci_blocks = [[(start, end, period, [seqs])]] per locus.
    Returns (AlignBatch, block_start, block_end)."""
    lbo, period, boo, stut, oso, starts, ends = [0], [], [0], [], [0], [], []
    oseq = bytearray()
    n_haps = 0
    for blocks in loci_blocks:
        H = 1
        for st, en, per, seqs in blocks:
            period.append(per)
            starts.append(st)
            ends.append(en)
            stut.extend(stutter)
            for q in seqs:
                oseq.extend(q.encode())
                oso.append(len(oseq))
            boo.append(len(oso) - 1)
            H *= len(seqs)
        lbo.append(len(period))
        n_haps += H
    arrs = dict(lbo=np.array(lbo, np.int32), period=np.array(period, np.int32), boo=np.array(boo, np.int32),
                stut=np.array(stut, np.float64), oso=np.array(oso, np.int32), oseq=bytes(oseq) + b"\0")
    b = AlignBatch()
    b.n_loci, b.n_blocks, b.n_options, b.n_pools, b.n_haps = len(loci_blocks), len(period), len(oso) - 1, 0, n_haps
    b.locus_block_off = ptr(arrs["lbo"], c_i32p)
    b.block_period = ptr(arrs["period"], c_i32p)
    b.block_opt_off = ptr(arrs["boo"], c_i32p)
    b.block_stutter = ptr(arrs["stut"], c_f64p)
    b.opt_seq_off = ptr(arrs["oso"], c_i32p)
    b.opt_seq = arrs["oseq"]
    b._keep = arrs
    return b, np.array(starts, np.int32), np.array(ends, np.int32)


def make_locus_reads(locus_read_off, locus_sample_off, reads, sample_label, name_id, log_p1, log_p2, haploid, rev_strand=None,
                     use_for_haps=None):
    """hipstr_locus_reads_t from Python lists: reads = [(start, stop, bases, quals, [(op, len)])] over all loci."""
    so, co = [0], [0]
    bases, quals, ctype, clen = bytearray(), bytearray(), bytearray(), []
    for start, stop, b, q, cig in reads:
        bases += b.encode()
        quals += q.encode()
        so.append(len(bases))
        for t, n in cig:
            ctype += t.encode()
            clen.append(n)
        co.append(len(clen))
    keep = dict(lro=np.ascontiguousarray(locus_read_off, np.int32), lso=np.ascontiguousarray(locus_sample_off, np.int32),
                so=np.array(so, np.int32), bases=np.frombuffer(bytes(bases) + b"\0", np.uint8).copy(),
                quals=np.frombuffer(bytes(quals) + b"\0", np.uint8).copy(), start=np.array([r[0] for r in reads], np.int32),
                stop=np.array([r[1] for r in reads], np.int32), co=np.array(co, np.int32),
                ctype=np.frombuffer(bytes(ctype) + b"\0", np.uint8).copy(), clen=np.array(clen + [0], np.int32),
                label=np.ascontiguousarray(sample_label, np.int32), name=np.ascontiguousarray(name_id, np.int32),
                p1=np.ascontiguousarray(log_p1, np.float64), p2=np.ascontiguousarray(log_p2, np.float64),
                hap=np.ascontiguousarray(haploid, np.uint8),
                rev=np.ascontiguousarray(rev_strand if rev_strand is not None else np.zeros(len(reads)), np.uint8),
                use=None if use_for_haps is None else np.ascontiguousarray(use_for_haps, np.uint8))
    rs = LocusReadsStruct(ptr(keep["lro"], c_i32p), ptr(keep["lso"], c_i32p), ptr(keep["so"], c_i32p), keep["bases"].ctypes.data,
                          keep["quals"].ctypes.data, ptr(keep["start"], c_i32p), ptr(keep["co"], c_i32p), keep["ctype"].ctypes.data,
                          ptr(keep["clen"], c_i32p), ptr(keep["label"], c_i32p), ptr(keep["name"], c_i32p), ptr(keep["p1"], c_f64p),
                          ptr(keep["p2"], c_f64p), ptr(keep["hap"], c_u8p), ptr(keep["rev"], c_u8p), ptr(keep["stop"], c_i32p),
                          ptr(keep["use"], c_u8p))
    rs._keep = keep
    return rs


def read_locus_reads(rs, n_loci):
    """The reads of a hipstr_locus_reads_t back as Python tuples [(start, stop, bases, quals, [(op, len)])] + locus_read_off."""
    lro = np.ctypeslib.as_array(rs.locus_read_off, shape=(n_loci + 1,)).copy()
    R = int(lro[-1])
    if R == 0:
        return [], lro
    so = np.ctypeslib.as_array(rs.read_seq_off, shape=(R + 1,))
    co = np.ctypeslib.as_array(rs.cigar_off, shape=(R + 1,))
    bases = C.string_at(rs.bases, int(so[-1]))
    quals = C.string_at(rs.quals, int(so[-1]))
    ctype = C.string_at(rs.cigar_type, int(co[-1]))
    clen = np.ctypeslib.as_array(rs.cigar_len, shape=(max(int(co[-1]), 1),))
    start = np.ctypeslib.as_array(rs.read_start, shape=(R,))
    stop = np.ctypeslib.as_array(rs.read_stop, shape=(R,))
    out = []
    for r in range(R):
        out.append((int(start[r]), int(stop[r]), bases[so[r]:so[r + 1]].decode(), quals[so[r]:so[r + 1]].decode(),
                    [(ctype[c:c + 1].decode(), int(clen[c])) for c in range(co[r], co[r + 1])]))
    return out, lro


class LeftAligned:
    """hipstr_left_aligned_t: the result of hipstr_left_align_reads_host (owns a hipstr_locus_reads_t)."""

    def __init__(self, ctx, n_loci, raw, chrom_seqs, trim_start=None, trim_stop=None):
        self.lib, self.n_loci = load(), n_loci
        chroms = [c if isinstance(c, bytes) else c.encode() for c in chrom_seqs]
        carr = (C.c_char_p * n_loci)(*chroms)
        ts = None if trim_start is None else np.ascontiguousarray(trim_start, np.int32)
        te = None if trim_stop is None else np.ascontiguousarray(trim_stop, np.int32)
        self._keep = (raw, chroms, carr, ts, te)
        h = C.c_void_p()
        st = self.lib.hipstr_left_align_reads_host(ctx.h if ctx else None, n_loci, C.byref(raw), carr, ptr(ts, c_i32p), ptr(te, c_i32p),
                                                   C.byref(h))
        if st != 0:
            raise HipstrError(st, "left_align_reads_host")
        self.h = h
        self.view = self.lib.hipstr_left_aligned_reads(h).contents
        n = C.c_int64()
        src = self.lib.hipstr_left_aligned_source(h, C.byref(n))
        self.source = np.ctypeslib.as_array(src, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int32)
        a, b = C.c_int64(), C.c_int64()
        self.lib.hipstr_left_aligned_counts(h, C.byref(a), C.byref(b))
        self.failed, self.nw_alignments = a.value, b.value

    def reads(self):
        return read_locus_reads(self.view, self.n_loci)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_left_aligned_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Genotyper:
    """hipstr_genotyper_t: a batch of loci run through the SeqStutterGenotyper::genotype() loop on the GPU."""

    def __init__(self, ctx, blocks, block_start, block_end, reads_struct, n_loci):
        """ctx = a Context, or None for host-only construction (genotype() then raises NO_DEVICE)."""
        self.lib, self.ctx, self.n_loci = load(), ctx, n_loci
        self._keep = (blocks, block_start, block_end, reads_struct)
        h = C.c_void_p()
        st = self.lib.hipstr_genotyper_create(ctx.h if ctx else None, C.byref(blocks), ptr(block_start, c_i32p),
                                              ptr(block_end, c_i32p), C.byref(reads_struct), C.byref(h))
        if st != 0:
            raise HipstrError(st, "genotyper_create")
        self.h = h

    @staticmethod
    def _reads_struct(synth):
        v = synth.view
        return LocusReadsStruct(v.locus_read_off, v.locus_sample_off, v.read_seq_off, v.read_bases, v.read_quals, v.read_start,
                                v.read_cigar_off, v.read_cigar_type, v.read_cigar_len, v.sample_label, v.read_name_id,
                                v.log_p1, v.log_p2, v.haploid, v.read_rev_strand, v.read_stop)

    @classmethod
    def from_synth_reads(cls, ctx, synth, stutter=(0.95, 0.05, 0.05, 0.95, 0.01, 0.01), loci_range=None, use_for_haps=None,
                         ref_alleles=None):
        """The full seam-B1 constructor: haplotype blocks are generated from the reads (hipstr_genotyper_create_from_reads).
        loci_range = (first, end) restricts the batch to a window of the Synth's loci."""
        v = synth.view
        l0, l1 = loci_range if loci_range else (0, synth.n_loci)
        L = l1 - l0
        rs = cls._reads_struct(synth)
        if use_for_haps is not None:   # [total reads] 0/1: which reads may propose candidate alleles
            use_for_haps = np.ascontiguousarray(use_for_haps, np.uint8)
            rs.use_for_haps = ptr(use_for_haps, c_u8p)
        if l0:   # the per-locus arrays start at the window; read-level arrays stay absolute
            step = C.sizeof(C.c_int32) * l0
            rs.locus_read_off = C.cast(C.cast(v.locus_read_off, C.c_void_p).value + step, c_i32p)
            rs.locus_sample_off = C.cast(C.cast(v.locus_sample_off, C.c_void_p).value + step, c_i32p)
            rs.haploid = C.cast(C.cast(v.haploid, C.c_void_p).value + l0, c_u8p)
        cl = int(v.chrom_len)
        raw = C.string_at(C.cast(v.chrom_seqs, C.c_void_p).value + l0 * cl, L * cl)
        chroms = [raw[l * cl:(l + 1) * cl] for l in range(L)]
        carr = (C.c_char_p * L)(*chroms)
        start = np.full(L, int(v.region_start), np.int32)
        stop = np.full(L, int(v.region_stop), np.int32)
        period = np.full(L, int(synth.cfg.period) or 4, np.int32)
        st6 = np.tile(np.asarray(stutter, np.float64), L)
        g = cls.__new__(cls)
        g.lib, g.ctx, g.n_loci = load(), ctx, L
        g._keep = (rs, chroms, carr, start, stop, period, st6, use_for_haps)
        g._synth = synth
        h = C.c_void_p()
        if ref_alleles is not None:   # [(pos or -1, [allele, ...])] per locus: the --ref-vcf reference panel
            apos = np.array([a[0] for a in ref_alleles], np.int32)
            aoff = np.cumsum([0] + [len(a[1]) for a in ref_alleles]).astype(np.int32)
            flat = [x.encode() for a in ref_alleles for x in a[1]]
            aarr = (C.c_char_p * max(len(flat), 1))(*flat)
            g._keep += (apos, aoff, flat, aarr)
            st = g.lib.hipstr_genotyper_create_with_ref_alleles(ctx.h if ctx else None, L, ptr(start, c_i32p), ptr(stop, c_i32p), ptr(period, c_i32p),
                                                                carr, ptr(st6, c_f64p), C.byref(rs), ptr(apos, c_i32p), ptr(aoff, c_i32p), aarr,
                                                                C.byref(h))
        else:
            st = g.lib.hipstr_genotyper_create_from_reads(ctx.h if ctx else None, L, ptr(start, c_i32p), ptr(stop, c_i32p),
                                                          ptr(period, c_i32p), carr, ptr(st6, c_f64p), C.byref(rs), C.byref(h))
        if st != 0:
            raise HipstrError(st, "genotyper_create_from_reads")
        g.h = h
        return g

    @classmethod
    def from_reads(cls, ctx, reads_struct, n_loci, region_start, region_stop, period, chrom_seqs,
                   stutter=(0.95, 0.05, 0.05, 0.95, 0.01, 0.01)):
        """hipstr_genotyper_create_from_reads on any hipstr_locus_reads_t (e.g. the output of LeftAligned)."""
        chroms = [c if isinstance(c, bytes) else c.encode() for c in chrom_seqs]
        carr = (C.c_char_p * n_loci)(*chroms)
        start, stop, per = (np.ascontiguousarray(a, np.int32) for a in (region_start, region_stop, period))
        st6 = np.asarray(stutter, np.float64)     # one model for every locus, or [n_loci][6]
        st6 = np.tile(st6, n_loci) if st6.ndim == 1 else np.ascontiguousarray(st6.reshape(-1))
        assert st6.size == 6 * n_loci
        g = cls.__new__(cls)
        g.lib, g.ctx, g.n_loci = load(), ctx, n_loci
        g._keep = (reads_struct, chroms, carr, start, stop, per, st6)
        h = C.c_void_p()
        st = g.lib.hipstr_genotyper_create_from_reads(ctx.h if ctx else None, n_loci, ptr(start, c_i32p), ptr(stop, c_i32p),
                                                      ptr(per, c_i32p), carr, ptr(st6, c_f64p), C.byref(reads_struct), C.byref(h))
        if st != 0:
            raise HipstrError(st, "genotyper_create_from_reads")
        g.h = h
        return g

    @classmethod
    def from_synth(cls, ctx, synth, loci_blocks=None):
        """All loci of a Synth; loci_blocks overrides the generator's own haplotype blocks."""
        v = synth.view
        rs = cls._reads_struct(synth)
        if loci_blocks is None:
            b = synth.batch
            bs = _np(v.block_start, b.n_blocks, np.int32)
            be = _np(v.block_end, b.n_blocks, np.int32)
        else:
            b, bs, be = blocks_batch(loci_blocks)
        g = cls(ctx, b, bs, be, rs, synth.n_loci)
        g._synth = synth
        return g

    def genotype(self, max_total_haplotypes=1000, max_flank_haplotypes=4, min_flank_freq=0.01, reassemble_flanks=False):
        ok = np.zeros(self.n_loci, np.uint8)
        st = self.lib.hipstr_genotyper_genotype(self.h, max_total_haplotypes, max_flank_haplotypes, min_flank_freq,
                                                int(reassemble_flanks), ptr(ok, c_u8p))
        if st != 0:
            raise HipstrError(st, "genotyper_genotype: " + (self.lib.hipstr_genotyper_last_error(self.h) or b"").decode())
        return ok

    def recompute_stutter_models(self, max_total_haplotypes=1000, max_flank_haplotypes=4, min_flank_freq=0.01, max_em_iter=100,
                                 abs_ll_converge=0.01, frac_ll_converge=0.001):
        ok = np.zeros(self.n_loci, np.uint8)
        st = self.lib.hipstr_genotyper_recompute_stutter_models(self.h, max_total_haplotypes, max_flank_haplotypes, min_flank_freq,
                                                                max_em_iter, abs_ll_converge, frac_ll_converge, ptr(ok, c_u8p))
        if st != 0:
            raise HipstrError(st, "recompute_stutter_models: " + (self.lib.hipstr_genotyper_last_error(self.h) or b"").decode())
        return ok

    def stats(self):
        a, t, r = C.c_int64(), C.c_int64(), C.c_int32()
        self.lib.hipstr_genotyper_stats(self.h, C.byref(a), C.byref(t), C.byref(r))
        return dict(alignments=a.value, traces=t.value, rounds=r.value)

    def timing(self):
        t = np.zeros(9)
        self.lib.hipstr_genotyper_timing(self.h, ptr(t, c_f64p))
        return dict(zip(("construct", "decide", "trace_device", "trace_host", "align", "posteriors", "vcf", "align_pack",
                         "align_unpack"), map(float, t)))

    def phase_timing(self):
        t = np.zeros(8)
        self.lib.hipstr_genotyper_phase_timing(self.h, ptr(t, c_f64p))
        return dict(zip(("align_all", "stutter", "prune_uncalled", "prune_unspanned", "assemble", "assemble_prune", "done", "failed"),
                        map(float, t)))

    def info(self, l):
        info = np.zeros(8, np.int32)
        self.lib.hipstr_genotyper_locus_info(self.h, l, ptr(info, c_i32p))
        return dict(zip(("blocks", "haps", "reads", "samples", "pools", "options", "seq_bytes", "rounds"), map(int, info)))

    def blocks(self, l):
        """[[sequences]] per block of locus l."""
        i = self.info(l)
        n = np.zeros(i["blocks"], np.int32)
        off = np.zeros(i["options"] + 1, np.int32)
        buf = np.zeros(max(i["seq_bytes"], 1), np.uint8)
        self.lib.hipstr_genotyper_locus_blocks(self.h, l, ptr(n, c_i32p), ptr(off, c_i32p), buf.ctypes.data)
        raw, out, o = bytes(buf), [], 0
        for k in n:
            out.append([raw[off[o + j]:off[o + j + 1]].decode() for j in range(k)])
            o += int(k)
        return out

    def results(self, l):
        i = self.info(l)
        R, S, H = i["reads"], i["samples"], i["haps"]
        o = dict(read_ll=np.zeros(R * H), seeds=np.zeros(R, np.int32), pool_index=np.zeros(R, np.int32),
                 post=np.zeros(S * H * H), sample_ll=np.zeros(S), best=np.zeros(S * 2, np.int32), call_ok=np.zeros(S, np.uint8))
        self.lib.hipstr_genotyper_locus_results(self.h, l, ptr(o["read_ll"], c_f64p), ptr(o["seeds"], c_i32p),
                                                ptr(o["pool_index"], c_i32p), ptr(o["post"], c_f64p),
                                                ptr(o["sample_ll"], c_f64p), ptr(o["best"], c_i32p), ptr(o["call_ok"], c_u8p))
        o["n_haps"] = H
        o["read_ll"] = o["read_ll"].reshape(R, H)
        o["post"] = o["post"].reshape(S, H, H)
        o["best"] = o["best"].reshape(S, 2)
        return o

    @staticmethod
    def vcf_loci(chroms, names, region_start, region_stop, period, chrom_seqs, locus_sample_names, out_sample_names):
        """Build a hipstr_vcf_loci_t (lists of str / bytes per locus; sample names flattened over loci)."""
        def arr(xs):
            xs = [x if isinstance(x, bytes) else x.encode() for x in xs]
            a = (C.c_char_p * len(xs))(*xs)
            a._keep = xs
            return a
        keep = [arr(chroms), arr(names), np.ascontiguousarray(region_start, np.int32), np.ascontiguousarray(region_stop, np.int32),
                np.ascontiguousarray(period, np.int32), arr(chrom_seqs), arr(locus_sample_names), arr(out_sample_names)]
        v = VcfLoci(keep[0], keep[1], ptr(keep[2], c_i32p), ptr(keep[3], c_i32p), ptr(keep[4], c_i32p), keep[5], keep[6],
                    len(out_sample_names), keep[7])
        v._keep = keep
        return v

    def write_vcf(self, loci, **options):
        """hipstr_genotyper_write_vcf; options override hipstr_vcf_default_options. Returns [(pos, text) or None] per locus."""
        opt = VcfOptions()
        self.lib.hipstr_vcf_default_options(C.byref(opt))
        for k, val in options.items():
            setattr(opt, k, val)
        st = self.lib.hipstr_genotyper_write_vcf(self.h, C.byref(loci), C.byref(opt))
        if st != 0:
            raise HipstrError(st, "genotyper_write_vcf: " + (self.lib.hipstr_genotyper_last_error(self.h) or b"").decode())
        out = []
        buf = np.zeros(1 << 22, np.uint8)
        for l in range(self.n_loci):
            pos = C.c_int32()
            n = self.lib.hipstr_genotyper_locus_record(self.h, l, C.byref(pos), buf.ctypes.data, len(buf))
            if n < 0:
                raise HipstrError(3, "record of locus %d needs %d bytes" % (l, -n))
            out.append((pos.value, bytes(buf[:n]).decode()) if n > 0 else None)
        return out

    def log(self, l):
        buf = np.zeros(1 << 16, np.uint8)
        n = self.lib.hipstr_genotyper_locus_log(self.h, l, buf.ctypes.data, len(buf))
        return bytes(buf[:max(n, 0)]).decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_genotyper_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchBuilder:
    """Hand-built batches for edge-case tests: add loci made of blocks and pooled reads."""

    def __init__(self):
        self.loci = []

    def add_locus(self, blocks, reads, stutter=(0.95, 0.05, 0.05, 0.95, 0.01, 0.01)):
        """blocks: list of (period, [option sequences]); reads: list of (bases, quals, seed)."""
        self.loci.append((blocks, reads, stutter))
        return self

    def build(self, realign_pool=None, realign_hap=None):
        lbo, lpo, lho, loo = [0], [0], [0], [0]
        period, boo, stut, oso, pso, seeds = [], [0], [], [0], [0], []
        oseq, pb, pq = bytearray(), bytearray(), bytearray()
        for blocks, reads, stutter in self.loci:
            H = 1
            for per, opts in blocks:
                period.append(per)
                stut.extend(stutter)
                for o in opts:
                    oseq.extend(o.encode())
                    oso.append(len(oseq))
                boo.append(len(oso) - 1)
                H *= len(opts)
            lbo.append(len(period))
            for bases, quals, seed in reads:
                assert len(bases) == len(quals)
                pb.extend(bases.encode())
                pq.extend(quals.encode() if isinstance(quals, str) else quals)
                pso.append(len(pb))
                seeds.append(seed)
            lpo.append(len(seeds))
            lho.append(lho[-1] + H)
            loo.append(loo[-1] + H * len(reads))
        keep = dict(
            lbo=np.array(lbo, np.int32), lpo=np.array(lpo, np.int32), lho=np.array(lho, np.int64),
            loo=np.array(loo, np.int64), period=np.array(period, np.int32), boo=np.array(boo, np.int32),
            stut=np.array(stut, np.float64), oso=np.array(oso, np.int32), pso=np.array(pso, np.int32),
            seeds=np.array(seeds, np.int32), oseq=bytes(oseq) + b"\0", pb=bytes(pb) + b"\0", pq=bytes(pq) + b"\0",
            rp=None if realign_pool is None else np.ascontiguousarray(realign_pool, np.uint8),
            rh=None if realign_hap is None else np.ascontiguousarray(realign_hap, np.uint8))
        b = AlignBatch()
        b.n_loci, b.n_blocks, b.n_options, b.n_pools, b.n_haps = len(self.loci), len(period), len(oso) - 1, len(seeds), lho[-1]
        b.locus_block_off, b.locus_pool_off = ptr(keep["lbo"], c_i32p), ptr(keep["lpo"], c_i32p)
        b.locus_hap_off, b.locus_out_off = ptr(keep["lho"], c_i64p), ptr(keep["loo"], c_i64p)
        b.block_period, b.block_opt_off = ptr(keep["period"], c_i32p), ptr(keep["boo"], c_i32p)
        b.block_stutter = ptr(keep["stut"], c_f64p)
        b.opt_seq_off, b.opt_seq = ptr(keep["oso"], c_i32p), keep["oseq"]
        b.pool_seq_off, b.pool_bases, b.pool_quals = ptr(keep["pso"], c_i32p), keep["pb"], keep["pq"]
        b.pool_seed = ptr(keep["seeds"], c_i32p)
        b.realign_pool, b.realign_hap = ptr(keep["rp"], c_u8p), ptr(keep["rh"], c_u8p)
        b._keep = keep
        b.n_out = int(loo[-1])
        return b


class MultiGenotyper:
    """hipstr_multi_t: the locus list of seam B1 dealt window by window to worker threads, one per (device, pipeline)."""

    STAGES = ("construct", "decide", "trace_device", "trace_host", "align", "posteriors", "vcf", "align_pack", "align_unpack")

    def __init__(self, devices=(0,), pipelines=2):
        self.lib = load()
        dev = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        st = self.lib.hipstr_multi_create(len(dev), ptr(dev, c_i32p), pipelines, C.byref(h))
        if st != 0:
            raise HipstrError(st, "hipstr_multi_create")
        self.h = h

    def close(self):
        if self.h:
            self.lib.hipstr_multi_destroy(self.h)
            self.h = None

    def window_order(self, synth, window_loci):
        n = self.lib.hipstr_multi_num_windows(synth.n_loci, window_loci)
        order = np.zeros(n, np.int32)
        st = self.lib.hipstr_multi_window_order(synth.n_loci, window_loci, synth.view.locus_read_off, ptr(order, c_i32p))
        if st != 0:
            raise HipstrError(st, "hipstr_multi_window_order")
        return order

    def genotype_synth(self, synth, vcf_loci, window_loci, stutter=(0.95, 0.05, 0.05, 0.95, 0.01, 0.01), next_window=None,
                       max_total_haplotypes=1000, max_flank_haplotypes=4, min_flank_freq=0.01, **options):
        """create_from_reads -> genotype(flank assembly on) -> write_vcf for every locus of a Synth.  next_window = a
        Python callable returning the next position in the dealing order (a dealer shared with other processes), or None
        for the handle's own counter.  Returns (locus_ok, [(pos, text) or None per locus])."""
        v = synth.view
        L = synth.n_loci
        rs = Genotyper._reads_struct(synth)
        cl = int(v.chrom_len)
        raw = C.string_at(v.chrom_seqs, L * cl)
        chroms = [raw[l * cl:(l + 1) * cl] for l in range(L)]
        carr = (C.c_char_p * L)(*chroms)
        start = np.full(L, int(v.region_start), np.int32)
        stop = np.full(L, int(v.region_stop), np.int32)
        period = np.full(L, int(synth.cfg.period) or 4, np.int32)
        st6 = np.tile(np.asarray(stutter, np.float64), L)
        opt = VcfOptions()
        self.lib.hipstr_vcf_default_options(C.byref(opt))
        for k, val in options.items():
            setattr(opt, k, val)
        ok = np.zeros(L, np.uint8)
        cb = NEXT_WINDOW_FN(lambda user: int(next_window())) if next_window is not None else C.cast(None, NEXT_WINDOW_FN)
        st = self.lib.hipstr_multi_genotype(self.h, L, ptr(start, c_i32p), ptr(stop, c_i32p), ptr(period, c_i32p), carr, ptr(st6, c_f64p),
                                            C.byref(rs), C.byref(vcf_loci), C.byref(opt), max_total_haplotypes, max_flank_haplotypes,
                                            min_flank_freq, window_loci, cb, None, ptr(ok, c_u8p))
        if st != 0:
            raise HipstrError(st, "hipstr_multi_genotype: " + (self.lib.hipstr_multi_last_error(self.h) or b"").decode())
        return ok, self.records(L)

    def records(self, n_loci):
        out = []
        buf = np.zeros(1 << 22, np.uint8)
        for l in range(n_loci):
            pos = C.c_int32()
            n = self.lib.hipstr_multi_locus_record(self.h, l, C.byref(pos), buf.ctypes.data, len(buf))
            if n < 0:
                raise HipstrError(3, "record of locus %d needs %d bytes" % (l, -n))
            out.append((pos.value, bytes(buf[:n]).decode()) if n > 0 else None)
        return out

    def stats(self):
        nw = self.lib.hipstr_multi_num_workers(self.h)
        a, t = C.c_int64(), C.c_int64()
        sec, win, busy = np.zeros(9), np.zeros(nw, np.int32), np.zeros(nw)
        self.lib.hipstr_multi_stats(self.h, C.byref(a), C.byref(t), ptr(sec, c_f64p), ptr(win, c_i32p), ptr(busy, c_f64p))
        h2d, d2h, nl = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.hipstr_multi_traffic(self.h, C.byref(h2d), C.byref(d2h), C.byref(nl))
        return {"alignments": a.value, "traces": t.value, "stage_seconds": dict(zip(self.STAGES, [round(float(x), 4) for x in sec])),
                "windows_per_worker": win.tolist(), "busy_seconds_per_worker": [round(float(x), 3) for x in busy],
                "h2d_bytes": h2d.value, "d2h_bytes": d2h.value, "gpu_launches": nl.value}


class Context:
    """hipstr_ctx_t on one GPU."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.hipstr_create(device, C.byref(h))
        if st != 0:
            raise HipstrError(st, "hipstr_create(device=%d)" % device)
        self.h = h

    def _check(self, st, what):
        if st != 0:
            raise HipstrError(st, "%s: %s" % (what, (self.lib.hipstr_last_error(self.h) or b"").decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.hipstr_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._check(self.lib.hipstr_set_stream(self.h, C.c_void_p(cuda_stream)), "set_stream")

    def align_host(self, batch, n_out, ll=None, want_pos=False):
        """hipstr_align_batch_host: host buffers in, host LLs out."""
        if ll is None:
            ll = np.zeros(n_out, np.float64)
        pos = np.full(n_out, -1, np.int32) if want_pos else None
        self._check(self.lib.hipstr_align_batch_host(self.h, C.byref(batch), ptr(ll, c_f64p), ptr(pos, c_i32p)),
                    "align_batch_host")
        return (ll, pos) if want_pos else ll

    def upload(self, batch):
        h = C.c_void_p()
        self._check(self.lib.hipstr_upload_batch(self.h, C.byref(batch), C.byref(h)), "upload_batch")
        return h

    def align_dev(self, handle, ll_dev_ptr, pos_dev_ptr=None):
        self._check(self.lib.hipstr_align_batch_dev(self.h, handle, C.c_void_p(ll_dev_ptr),
                                                    C.c_void_p(pos_dev_ptr) if pos_dev_ptr else None),
                    "align_batch_dev")

    def free_batch(self, handle):
        self.lib.hipstr_free_batch(self.h, handle)

    def genotype_host(self, batch, reads, n_elems, n_reads, post_size, n_samples, n_loci, read_ll=None,
                      read_seed=None, post=None, sample_ll=None, best=None, total_ll=None):
        """hipstr_genotype_batch_host: K1 + K2 + K3 with host buffers in and out (caller-owned, ideally page-locked,
        result buffers may be passed in; missing ones are allocated)."""
        out = dict(read_ll=np.zeros(n_elems, np.float64) if read_ll is None else read_ll,
                   read_seed=np.full(n_reads, -2, np.int32) if read_seed is None else read_seed,
                   post=np.zeros(post_size, np.float64) if post is None else post,
                   sample_ll=np.zeros(n_samples, np.float64) if sample_ll is None else sample_ll,
                   best=np.zeros(2 * n_samples, np.int32) if best is None else best.reshape(-1),
                   total_ll=np.zeros(n_loci, np.float64) if total_ll is None else total_ll)
        go = GenotypeOut(*[out[k].ctypes.data for k in ("read_ll", "read_seed", "post", "sample_ll", "best", "total_ll")])
        self._check(self.lib.hipstr_genotype_batch_host(self.h, C.byref(batch), C.byref(reads), C.byref(go)),
                    "genotype_batch_host")
        out["best"] = out["best"].reshape(-1, 2)
        return out

    def upload_genotype(self, batch, reads):
        h = C.c_void_p()
        self._check(self.lib.hipstr_upload_genotype_batch(self.h, C.byref(batch), C.byref(reads), C.byref(h)),
                    "upload_genotype_batch")
        return h

    def genotype_dev(self, handle, read_ll, read_seed, post, sample_ll, best, total_ll):
        """Device pointers (ints); read_seed / best / total_ll may be 0."""
        go = GenotypeOut(read_ll, read_seed or None, post, sample_ll, best or None, total_ll or None)
        self._check(self.lib.hipstr_genotype_batch_dev(self.h, handle, C.byref(go)), "genotype_batch_dev")

    def free_genotype(self, handle):
        self.lib.hipstr_free_genotype_batch(self.h, handle)

    def extract_genotypes(self, locus_sample_off, n_haps, n_variants, hap_to_allele, haploid, post, sample_ll):
        """hipstr_extract_genotypes_host -> dict of per-sample genotype outputs."""
        st, out = extract_genotypes(self.lib.hipstr_extract_genotypes_host, locus_sample_off, n_haps, n_variants,
                                    hap_to_allele, haploid, post, sample_ll, self.h)
        self._check(st, "extract_genotypes_host")
        return out

    def trace(self, batch, block_start, trace_pool, trace_hap, aln_stride=1024):
        """hipstr_trace_batch_host -> dict of traceback outputs (see hipstr_trace_out_t)."""
        st, out = trace_batch(self.lib.hipstr_trace_batch_host, batch, block_start, trace_pool, trace_hap, aln_stride, self.h)
        self._check(st, "trace_batch_host")
        return out

    def nw_align(self, refs, reads, use_ref_end_penalty=False):
        """hipstr_nw_align_batch_host on lists of str -> ([ops], scores float32)."""
        n = len(refs)
        ro = np.zeros(n + 1, np.int32)
        qo = np.zeros(n + 1, np.int32)
        ro[1:] = np.cumsum([len(r) for r in refs])
        qo[1:] = np.cumsum([len(r) for r in reads])
        stride = max(len(a) for a in refs) + max(len(b) for b in reads) + 2
        ops = np.zeros(n * stride, np.uint8)
        lens = np.zeros(n, np.int32)
        score = np.zeros(n, np.float32)
        self._check(self.lib.hipstr_nw_align_batch_host(self.h, n, ptr(ro, c_i32p), "".join(refs).encode(), ptr(qo, c_i32p),
                                                        "".join(reads).encode(), int(use_ref_end_penalty), stride, ops.ctypes.data,
                                                        ptr(lens, c_i32p), score.ctypes.data_as(C.POINTER(C.c_float))),
                    "nw_align_batch_host")
        raw = ops.reshape(n, stride)
        return [bytes(raw[i, :max(int(lens[i]), 0)]).decode() if lens[i] >= 0 else None for i in range(n)], score

    def snp_phasing(self, batch):
        """hipstr_snp_phasing_batch_host on a SnpPhasing -> (log_p1, log_p2, counts [n][4])."""
        st, p1, p2, counts = batch.run(self.lib.hipstr_snp_phasing_batch_host, self.h)
        self._check(st, "snp_phasing_batch_host")
        return p1, p2, counts

    def em_train(self, batch, max_iter=100, min_abs=0.01, min_frac=0.001):
        """hipstr_em_train_host -> (params [L][6], converged [L], iterations [L], final LL [L])."""
        st, prm, conv, it, ll = em_train(self.lib.hipstr_em_train_host, batch, max_iter, min_abs, min_frac, self.h)
        self._check(st, "em_train_host")
        return prm, conv, it, ll

    def trace_seconds(self):
        t = np.zeros(4)
        self.lib.hipstr_trace_seconds(self.h, ptr(t, c_f64p))
        return dict(zip(("lower", "order_upload", "kernel", "download"), map(float, t)))

    def traffic(self):
        a, b, n = C.c_int64(), C.c_int64(), C.c_int32()
        self.lib.hipstr_last_traffic(self.h, C.byref(a), C.byref(b), C.byref(n))
        return a.value, b.value, n.value

    def enable_timing(self, on=True):
        self._check(self.lib.hipstr_enable_timing(self.h, 1 if on else 0), "enable_timing")

    def collect_timing(self):
        """(K1 ms, scatter+posterior ms, calls) summed since the last collect; synchronises the stream."""
        a, b, n = C.c_double(), C.c_double(), C.c_int32()
        self._check(self.lib.hipstr_collect_timing(self.h, C.byref(a), C.byref(b), C.byref(n)), "collect_timing")
        return a.value, b.value, n.value

    def scatter_host(self, n_haps, pool_ll, pool_seed, pool_index, second_mate, read_ll, read_seed=None,
                     copy_read=None, realign_hap=None):
        n_reads = len(pool_index)
        self._check(self.lib.hipstr_scatter_pool_lls_host(
            self.h, n_reads, n_haps, ptr(pool_ll, c_f64p), ptr(pool_seed, c_i32p), ptr(pool_index, c_i32p),
            ptr(second_mate, c_u8p), ptr(copy_read, c_u8p), ptr(realign_hap, c_u8p), ptr(read_ll, c_f64p),
            ptr(read_seed, c_i32p)), "scatter_pool_lls_host")
        return read_ll

    def posteriors_host(self, locus_read_off, locus_sample_off, n_haps, haploid, read_ll, log_p1, log_p2,
                        sample_label, read_weight):
        n_loci = len(n_haps)
        S = int(locus_sample_off[-1])
        post_size = int(sum(int(locus_sample_off[l + 1] - locus_sample_off[l]) * int(n_haps[l]) ** 2
                            for l in range(n_loci)))
        post = np.zeros(post_size, np.float64)
        sample_ll = np.zeros(S, np.float64)
        best = np.zeros(2 * S, np.int32)
        total = np.zeros(n_loci, np.float64)
        self._check(self.lib.hipstr_posteriors_host(
            self.h, n_loci, ptr(locus_read_off, c_i32p), ptr(locus_sample_off, c_i32p), ptr(n_haps, c_i32p),
            ptr(haploid, c_u8p), ptr(read_ll, c_f64p), ptr(log_p1, c_f64p), ptr(log_p2, c_f64p),
            ptr(sample_label, c_i32p), ptr(read_weight, c_i32p), ptr(post, c_f64p), ptr(sample_ll, c_f64p),
            ptr(best, c_i32p), ptr(total, c_f64p)), "posteriors_host")
        return post, sample_ll, best.reshape(-1, 2), total
