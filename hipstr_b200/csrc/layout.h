/*
 * layout.h -- device-resident layout of a flattened batch of loci.
 *
 * The reference keeps a locus as a graph of C++ objects (Haplotype -> HapBlock /
 * RepeatBlock -> StutterAlignerClass, std::vector<Alignment>); here the host
 * flattens a whole batch of loci ONCE into a handful of packed arrays that are
 * uploaded with one cudaMemcpyAsync each and stay resident in HBM:
 *
 *   pools      one record per pooled read: 16B-aligned offset of its bases/quals
 *              (so a pool is fetched into shared memory with one bulk copy), length,
 *              seed, owning locus, output row
 *   hapsides   one record per (locus, haplotype, orientation): oriented sequence,
 *              per-row homopolymer class, block list.  Orientation 0 = forward
 *              haplotype (left of the seed), 1 = reversed haplotype (right of the
 *              seed), as in HapAligner.cpp:606-627.
 *   reps       one record per (repeat-block allele, orientation): the position walks the
 *              reference derives from StutterAlignerClass's upstream_match_lengths_
 *              (StutterAlignerClass.h:35-80), unrolled into programs, + the 13 PCR-artifact
 *              log-priors (RepeatStutterInfo.h:53-61), computed on the host with glibc.
 *   jobs       (pool, haplotype range) work items, bucketed by columns-per-lane.
 */
#ifndef HIPSTR_B200_LAYOUT_H_
#define HIPSTR_B200_LAYOUT_H_

#include <stdint.h>
#ifdef __CUDACC__
#define HIPSTR_HD __host__ __device__
#else
#define HIPSTR_HD
#endif

#define HIPSTR_WARPS_PER_CTA 1
#define HIPSTR_MAX_BLOCKS 8          /* haplotype blocks per locus handled by the kernel */
#define HIPSTR_ROW_REPEAT 0x80       /* rowinfo flag: row belongs to a repeat block      */
#define HIPSTR_ROW_AFTER_REPEAT 0x40 /* rowinfo flag: first row after a repeat block     */

struct DevPool {          /* 32 B */
  int32_t seq_off;        /* byte offset into bases/quals, multiple of 16 */
  int32_t len;            /* read length */
  int32_t seed;           /* seed base index (>= 1, <= len-2); never -1 here */
  int32_t locus;
  int64_t out_off;        /* ll_out index of haplotype 0 of this pool */
  int32_t hap_rec0;       /* first hapside record of the locus = 2 * (global hap index of hap 0) */
  int32_t n_haps;
};

struct DevHapSide {       /* 32 B */
  int32_t seq_off;        /* into hapbytes: oriented sequence, len bytes */
  int32_t row_off;        /* into hapbytes: rowinfo, len bytes: low 4 bits = homopolymer class
                             (min(15, max(hp(i), hp(i-1)))), flags above */
  int32_t len;            /* haplotype length L_h */
  int32_t blk_off;        /* into blocks */
  int32_t n_blocks;
  int32_t n_seed_pos;     /* total length of flank blocks (compute_aln_logprob num_seeds) */
  int32_t seg1_class;     /* >= 0: haplotypes of a locus with equal class share every row before the
                             first repeat block in this orientation (sequence AND homopolymer
                             classes), so the kernel reuses those rows from the previous haplotype
                             of the job; -1: never reuse (non flank/repeat/flank structures) */
  int32_t first_rep;      /* index of the first repeat block in this orientation (n_blocks if none) */
};

struct DevBlock {         /* 16 B */
  int32_t row_start;      /* first haplotype row of the block in this orientation */
  int32_t len;
  int32_t rep;            /* index into reps, or -1 for a flank block */
  int32_t tslot;          /* repeat block: slot of this (block, allele) in the stutter tables of its locus (K1a) */
};

/* One step of a repeat-block "program": the position walk of align_pcr_insertion_reverse /
 * align_pcr_deletion_reverse (StutterAlignerClass.cpp:75-100,127-147) depends only on the allele
 * (its upstream_match_lengths_ table), not on the read, so the host unrolls it once per allele and
 * every read column replays it: no data-dependent pointer chasing on the device, and the next
 * step can be prefetched while the current one is applied.
 *
 * A step is one of two things, told apart by `moves` (uniform across the warp that replays the walk):
 *     moves != 0:  lp = (lp - val[col + off_a]) + val[col + off_b];   term = lp
 *     moves == 0:  term = lp + logrun      -- a run of positions that all share lp, collapsed by the reference into one
 *                                             term with int_log(run length) added; logrun = 0.0 where nothing is added
 * where val is the per-read emission table with HIPSTR_VAL_STRIDE doubles per read column (base codes 0..4).  A step
 * never does both, so the two offsets and the double share the same 8 bytes and a step is ONE 16-byte load.  Offsets
 * are BYTES relative to the table entry of the column the walk started at (insertion walks: relative to column
 * j - period, the first base an inserted copy overwrites). */
#define HIPSTR_VAL_STRIDE 5   /* odd stride in 8-byte words: consecutive columns fall in distinct banks */
struct DevProgEntry {     /* 16 B */
  int32_t pos;            /* artifact position i (<= 0, offset from the right end of the block); the
                             terminal entry of a walk holds the position where the walk stops */
  int32_t moves;          /* 1 if the step changes lp (insertions repeat it once per inserted copy) */
  union {
    struct { int32_t off_a, off_b; };   /* moves != 0: byte offsets into val of the emission to remove / to add */
    double logrun;                      /* moves == 0: added to lp for this step's term */
  };
};

struct DevRep {           /* 168 B */
  int32_t seq_off;        /* into hapbytes: oriented allele base codes */
  int32_t len;            /* B */
  int32_t period;
  int32_t n_del;          /* StutterAlignerClass num_deletions_ */
  int32_t left_align;     /* !reversed (RepeatBlock.h:28,41); only the traceback uses it */
  int32_t prog_off[7];    /* into progs: [0] insertion walk (lag = period), [k] deletion of k units */
  int32_t diag_off;       /* into rep_tabs: B byte offsets of the right-anchored diagonal,
                             entry t = emission of column q - t against allele base B-1-t, relative to column q */
  int32_t ins_off;        /* into rep_tabs: 6*period byte offsets of the periodic-copy sum
                             (StutterAlignerClass.cpp:38-51); -1 marks "log_correct of the read base" */
  double  art[13];        /* log_prob_pcr_artifact for D = -6p .. +6p */
};

/* K1a work item: the stutter tables of one pooled read against a range of (repeat block, allele) slots of its locus.
 * A slot pairs the forward-oriented DevRep (read bases left of the seed) with the reversed one (right of the seed). */
struct DevStutJob {       /* 16 B */
  int32_t pool;
  int32_t slot0;          /* first entry of slot_reps */
  int32_t n_slots;
  int32_t tslot0;         /* table slot of slot0 within the pool's slab of the stutter tables */
};
struct DevSlotReps { int32_t rep_fwd, rep_rev; };

/* Stutter tables T (K1a -> K1b), per pooled read a slab at pool_t_off[pool] (in doubles):
 *     T[tslot][artifact 0..12][column g],  g = side-order column (left of the seed 0..nL-1, then right nL..n-2),
 * row pitch = hipstr_t_pitch(len).  T = log_prob_pcr_artifact + align_stutter_region_reverse for the read prefix
 * ending at column g (HapAligner.cpp:84-87 without pre_prob); it depends only on (read, allele), not on the flanks. */
#define HIPSTR_STUT_SLOTS_PER_JOB 8
static inline HIPSTR_HD int32_t hipstr_t_pitch(int32_t read_len) { return (read_len + 1) & ~1; }

struct DevJob {           /* 16 B */
  int32_t pool;
  int32_t h0, h1;         /* haplotype range [h0, h1) of the pool's locus */
  int32_t pad;
};

struct AlignParams {
  const DevJob* jobs;
  int32_t n_jobs;
  int32_t n_max;          /* shared-memory stride: max read length of the launch (multiple of 16) */
  int32_t l_max;          /* max haplotype length of the launch (multiple of 2) */
  const DevPool* pools;
  const char* bases;
  const char* quals;
  const DevHapSide* hapsides;
  const uint8_t* hapbytes;
  const DevBlock* blocks;
  const DevRep* reps;
  const DevProgEntry* progs;
  const int32_t* rep_tabs;
  const uint8_t* hap_mask;   /* per global hap index; NULL = all */
  const double* qual_lut;    /* [256][2]: log_correct, log_error by quality byte */
  const double* trans;       /* [3][16]: LOG_MATCH_TO_MATCH / _INS / _DEL by homopolymer class */
  const double* int_logs;    /* [10000] */
  double* ll_out;
  int32_t* pos_out;          /* may be NULL */
  double* debug_out;         /* may be NULL: [2][l_max] last-column M values of job 0's last haplotype */
  int32_t* job_counter;      /* zeroed before the launch: persistent warps pull jobs from it */
  double* last_scratch;      /* [resident warps][2][l_max] last-column slabs */
  const double* stut;        /* stutter tables of this launch's pools (K1a output) */
  const int64_t* pool_t_off; /* [n_pools] offset of the pool's slab in stut, in doubles */
  /* ---- traces (K5 forward pass, k_align<.., TRACE = true>): one job per trace = (pool, ONE haplotype), job.pad = trace ---- */
  const int64_t* job_t_off;  /* [n_jobs of the launch] slab of the trace in stut / stut_pos (slot = ordinal of the repeat block) */
  const int32_t* stut_pos;   /* best artifact position of every table entry (K1a, TRACE) */
  unsigned char* dec;        /* predecessor choices, one byte per flank cell: left side [hap rows][nL] then right side [hap rows][nR] */
  const int64_t* dec_off;    /* [n_traces] */
  int32_t* art;              /* per trace: best artifact size [2 sides][blocks][n_side], then position, same shape */
  const int64_t* art_off;    /* [n_traces] */
  int32_t* trace_seed_pos;   /* [n_traces] haplotype position of the seed base */
};

struct StutParams {          /* K1a */
  const DevStutJob* jobs;
  int32_t n_jobs;
  int32_t n_max;             /* longest read of the launch, multiple of 16 */
  const DevPool* pools;
  const char* bases;
  const char* quals;
  const DevSlotReps* slot_reps;
  const DevRep* reps;
  const DevProgEntry* progs;
  const int32_t* rep_tabs;
  const double* qual_lut;
  const double* int_logs;
  const int64_t* pool_t_off;
  double* stut;
  int32_t* job_counter;
  /* traces (K5): one table slab per JOB instead of per pool, and the best artifact position of every walk */
  const int64_t* job_t_off;  /* NULL for alignment */
  int32_t* stut_pos;         /* NULL for alignment; parallel to stut */
};

#endif
