"""TEST INFRASTRUCTURE: ctypes view of oracle/ref_genotyper_harness.cpp -- the UNMODIFIED reference
SeqStutterGenotyper (seam B1) driven on one synthetic locus at a time."""
import ctypes as C

import numpy as np

import checkers
from hipstr_b200.capi import c_f64p, c_i32p, c_u8p, ptr, _np

DEF_STUTTER = (0.95, 0.05, 0.05, 0.95, 0.01, 0.01)


def bind(lib):
    if getattr(lib, "_sg_bound", False):
        return lib
    lib.ref_sg_create.restype = C.c_void_p
    lib.ref_sg_create.argtypes = [C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_void_p, C.c_void_p, c_i32p,
                                  C.c_void_p, c_i32p, c_f64p, c_f64p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32,
                                  c_f64p, C.c_int32, C.c_int32, c_u8p, c_u8p, C.c_char_p]
    lib.ref_sg_set_output_flags.restype = None
    lib.ref_sg_set_output_flags.argtypes = [c_i32p, C.c_double]
    for name in ("destroy",):
        getattr(lib, "ref_sg_" + name).restype = None
        getattr(lib, "ref_sg_" + name).argtypes = [C.c_void_p]
    for name in ("initialized", "num_blocks", "num_haps", "num_pools"):
        getattr(lib, "ref_sg_" + name).restype = C.c_int32
        getattr(lib, "ref_sg_" + name).argtypes = [C.c_void_p]
    lib.ref_sg_block_info.restype = None
    lib.ref_sg_block_info.argtypes = [C.c_void_p, C.c_int32, c_i32p]
    lib.ref_sg_block_seqs.restype = None
    lib.ref_sg_block_seqs.argtypes = [C.c_void_p, C.c_int32, c_i32p, C.c_void_p]
    lib.ref_sg_genotype.restype = C.c_int32
    lib.ref_sg_genotype.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double]
    lib.ref_sg_recompute_stutter_models.restype = C.c_int32
    lib.ref_sg_recompute_stutter_models.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double]
    lib.ref_sg_stutter_params.restype = None
    lib.ref_sg_stutter_params.argtypes = [C.c_void_p, c_f64p]
    lib.ref_sg_results.restype = None
    lib.ref_sg_results.argtypes = [C.c_void_p, c_f64p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_u8p]
    lib.ref_sg_write_vcf.restype = C.c_int32
    lib.ref_sg_write_vcf.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.ref_sg_log.restype = C.c_int32
    lib.ref_sg_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib._sg_bound = True
    return lib


class LocusReads:
    """The un-pooled reads of one synthetic locus (slices of the generator's arrays)."""

    def __init__(self, synth, l):
        v = synth.view
        R = int(v.n_reads)
        lro = _np(v.locus_read_off, synth.n_loci + 1, np.int32)
        r0, r1 = int(lro[l]), int(lro[l + 1])
        so = _np(v.read_seq_off, R + 1, np.int32)
        co = _np(v.read_cigar_off, R + 1, np.int32)
        nb, nc = int(so[R]), int(co[R])
        bases = np.ctypeslib.as_array(C.cast(v.read_bases, C.POINTER(C.c_uint8)), shape=(nb,))
        quals = np.ctypeslib.as_array(C.cast(v.read_quals, C.POINTER(C.c_uint8)), shape=(nb,))
        ctype = np.ctypeslib.as_array(C.cast(v.read_cigar_type, C.POINTER(C.c_uint8)), shape=(nc,))
        clen = _np(v.read_cigar_len, nc, np.int32)
        self.n_reads = r1 - r0
        self.n_samples = int(synth.locus_sample_off[l + 1] - synth.locus_sample_off[l])
        self.seq_off = (so[r0:r1 + 1] - so[r0]).astype(np.int32)
        self.bases = bases[so[r0]:so[r1]].copy()
        self.quals = quals[so[r0]:so[r1]].copy()
        self.cigar_off = (co[r0:r1 + 1] - co[r0]).astype(np.int32)
        self.cigar_type = ctype[co[r0]:co[r1]].copy()
        self.cigar_len = clen[co[r0]:co[r1]].copy()
        self.start = _np(v.read_start, R, np.int32)[r0:r1].copy()
        self.stop = _np(v.read_stop, R, np.int32)[r0:r1].copy()
        self.name_id = _np(v.read_name_id, R, np.int32)[r0:r1].copy()
        self.sample_label = synth.sample_label[r0:r1].copy()
        self.log_p1 = synth.log_p1[r0:r1].copy()
        self.log_p2 = synth.log_p2[r0:r1].copy()
        self.second_mate = synth.second_mate[r0:r1].copy()
        self.rev_strand = _np(v.read_rev_strand, R, np.uint8)[r0:r1].copy()
        cl = int(v.chrom_len)
        chrom = np.ctypeslib.as_array(C.cast(v.chrom_seqs, C.POINTER(C.c_uint8)), shape=(synth.n_loci * cl,))
        self.chrom_seq = bytes(chrom[l * cl:(l + 1) * cl])
        self.region = (int(v.region_start), int(v.region_stop))
        self.period = int(synth.cfg.period) or 4
        self.haploid = int(synth.haploid[l])


class ReadsOfLocus:
    """LocusReads built from Python tuples [(start, stop, bases, quals, [(op, len)])] (e.g. left-aligned reads)."""

    def __init__(self, reads, n_samples, sample_label, name_id, log_p1, log_p2, chrom_seq, region, period, haploid=0, rev_strand=None,
                 use_for_haps=None):
        self.n_reads, self.n_samples = len(reads), n_samples
        so, co, bases, quals, ctype, clen = [0], [0], bytearray(), bytearray(), bytearray(), []
        for start, stop, b, q, cig in reads:
            bases += b.encode()
            quals += q.encode()
            so.append(len(bases))
            for t, n in cig:
                ctype += t.encode()
                clen.append(n)
            co.append(len(clen))
        self.seq_off, self.cigar_off = np.array(so, np.int32), np.array(co, np.int32)
        self.bases = np.frombuffer(bytes(bases) + b"\0", np.uint8).copy()
        self.quals = np.frombuffer(bytes(quals) + b"\0", np.uint8).copy()
        self.cigar_type = np.frombuffer(bytes(ctype) + b"\0", np.uint8).copy()
        self.cigar_len = np.array(clen + [0], np.int32)
        self.start = np.array([r[0] for r in reads], np.int32)
        self.stop = np.array([r[1] for r in reads], np.int32)
        self.sample_label = np.ascontiguousarray(sample_label, np.int32)
        self.name_id = np.ascontiguousarray(name_id, np.int32)
        self.log_p1, self.log_p2 = np.ascontiguousarray(log_p1, np.float64), np.ascontiguousarray(log_p2, np.float64)
        self.rev_strand = np.ascontiguousarray(rev_strand if rev_strand is not None else np.zeros(len(reads)), np.uint8)
        self.chrom_seq, self.region, self.period, self.haploid = chrom_seq, region, period, haploid
        self.use_for_haps = None if use_for_haps is None else np.ascontiguousarray(use_for_haps, np.uint8)


class RefGenotyper:
    """One reference SeqStutterGenotyper object."""

    def __init__(self, reads, stutter=DEF_STUTTER, reassemble_flanks=False, ref_vcf=None):
        self.lib = bind(checkers.ref())
        self.reads = reads
        st = np.asarray(stutter, np.float64)
        self.h = self.lib.ref_sg_create(
            reads.n_samples, reads.n_reads, ptr(reads.sample_label, c_i32p), ptr(reads.name_id, c_i32p),
            ptr(reads.start, c_i32p), ptr(reads.stop, c_i32p), ptr(reads.seq_off, c_i32p), reads.bases.ctypes.data, reads.quals.ctypes.data,
            ptr(reads.cigar_off, c_i32p), reads.cigar_type.ctypes.data, ptr(reads.cigar_len, c_i32p),
            ptr(reads.log_p1, c_f64p), ptr(reads.log_p2, c_f64p), reads.chrom_seq, reads.region[0], reads.region[1],
            reads.period, ptr(st, c_f64p), reads.haploid, int(reassemble_flanks), ptr(reads.rev_strand, c_u8p),
            ptr(getattr(reads, "use_for_haps", None), c_u8p), ref_vcf.encode() if ref_vcf else None)
        self.initialized = bool(self.lib.ref_sg_initialized(self.h))

    def blocks(self):
        """[(start, end, period, [sequences])] of the current haplotype blocks."""
        out = []
        for b in range(self.lib.ref_sg_num_blocks(self.h)):
            info = np.zeros(5, np.int32)
            self.lib.ref_sg_block_info(self.h, b, ptr(info, c_i32p))
            off = np.zeros(info[3] + 1, np.int32)
            buf = np.zeros(max(int(info[4]), 1), np.uint8)
            self.lib.ref_sg_block_seqs(self.h, b, ptr(off, c_i32p), buf.ctypes.data)
            raw = bytes(buf)
            out.append((int(info[0]), int(info[1]), int(info[2]), [raw[off[i]:off[i + 1]].decode() for i in range(info[3])]))
        return out

    def genotype(self, max_total_haps=1000, max_flank_haps=4, min_flank_freq=0.01):
        return bool(self.lib.ref_sg_genotype(self.h, max_total_haps, max_flank_haps, min_flank_freq))

    def recompute_stutter_models(self, max_total_haps=1000, max_flank_haps=4, min_flank_freq=0.01, max_em_iter=100, abs_ll=0.01,
                                 frac_ll=0.001):
        return bool(self.lib.ref_sg_recompute_stutter_models(self.h, max_total_haps, max_flank_haps, min_flank_freq, max_em_iter,
                                                             abs_ll, frac_ll))

    def stutter_params(self):
        out = np.zeros(6)
        self.lib.ref_sg_stutter_params(self.h, ptr(out, c_f64p))
        return out

    def results(self):
        R, S, H = self.reads.n_reads, self.reads.n_samples, self.lib.ref_sg_num_haps(self.h)
        o = dict(read_ll=np.zeros(R * H), seeds=np.zeros(R, np.int32), pool_index=np.zeros(R, np.int32),
                 post=np.zeros(S * H * H), sample_ll=np.zeros(S), best=np.zeros(S * 2, np.int32),
                 call_ok=np.zeros(S, np.uint8))
        self.lib.ref_sg_results(self.h, ptr(o["read_ll"], c_f64p), ptr(o["seeds"], c_i32p), ptr(o["pool_index"], c_i32p),
                                ptr(o["post"], c_f64p), ptr(o["sample_ll"], c_f64p), ptr(o["best"], c_i32p),
                                ptr(o["call_ok"], c_u8p))
        o["n_haps"] = H
        o["read_ll"] = o["read_ll"].reshape(R, H)
        o["post"] = o["post"].reshape(S, H, H)
        o["best"] = o["best"].reshape(S, 2)
        return o

    def vcf(self, output_gls=0, output_pls=0, output_phased_gls=0, output_allreads=1, output_mallreads=1, output_filters=0,
            output_haplotype_data=0, max_flank_indel_frac=0.15):
        flags = np.array([output_gls, output_pls, output_phased_gls, output_allreads, output_mallreads, output_filters,
                          output_haplotype_data], np.int32)
        self.lib.ref_sg_set_output_flags(ptr(flags, c_i32p), max_flank_indel_frac)
        buf = np.zeros(1 << 22, np.uint8)
        n = self.lib.ref_sg_write_vcf(self.h, buf.ctypes.data, len(buf))
        assert n >= 0
        return bytes(buf[:n]).decode()

    def log(self):
        buf = np.zeros(1 << 20, np.uint8)
        n = self.lib.ref_sg_log(self.h, buf.ctypes.data, len(buf))
        return bytes(buf[:n]).decode()

    def close(self):
        if self.h:
            self.lib.ref_sg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
