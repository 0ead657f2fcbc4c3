/*
 * hipstr_b200.h -- C-ABI of the B200-native HipSTR hot path.
 *
 * Every entry point replaces one seam of the reference (tfwillems/HipSTR); the
 * reference interface it stands in for is cited next to each declaration as
 * path:line relative to the reference checkout.  The reference has no FFI of
 * its own (it is one C++ program), so the "binding" a maintainer adds is a
 * direct C++ call from the class that owns the seam -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; all functions return a
 *     hipstr_status_t (0 = OK) instead of exit()/assert() like the reference;
 *   - "host" entry points take HOST buffers and do the H2D/D2H copies
 *     themselves (end-to-end path); "_dev" entry points take DEVICE buffers
 *     (inputs resident in HBM) and only enqueue kernels on the given stream;
 *   - a context owns one GPU's streams, workspaces and the constant tables
 *     (quality LUT, transition tables, INT_LOGS) which are computed ON THE HOST
 *     with glibc and uploaded -- never recomputed with CUDA libm;
 *   - no CPU fallback: if no CUDA device is usable hipstr_create() fails with
 *     HIPSTR_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef HIPSTR_B200_H_
#define HIPSTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HIPSTR_OK               = 0,
  HIPSTR_ERR_NO_DEVICE    = 1,  /* no usable CUDA device / driver            */
  HIPSTR_ERR_CUDA         = 2,  /* a CUDA runtime call failed (see last_error) */
  HIPSTR_ERR_BAD_ARG      = 3,  /* malformed batch (offsets, sizes, NULLs)   */
  HIPSTR_ERR_UNSUPPORTED  = 4,  /* shape outside the kernel's static limits  */
  HIPSTR_ERR_INVALID_SEED = 5,  /* reference: printErrorAndDie("Invalid alignment seed") */
  HIPSTR_ERR_BAD_CIGAR    = 6   /* reference: "Unrecognized CIGAR char in calc_seed_base()" */
} hipstr_status_t;

/* Number of PCR-stutter artifact sizes evaluated per repeat block:
 * -6..+6 repeat units (RepeatStutterInfo.h:10-11). */
#define HIPSTR_NUM_ARTIFACTS 13
#define HIPSTR_MAX_ARTIFACT_UNITS 6

/* ------------------------------------------------------------------------- *
 *  Flat description of a batch of loci for the alignment HMM.
 *
 *  One locus = the candidate-haplotype structure the reference keeps in
 *  Haplotype/HapBlock/RepeatBlock objects (SeqAlignment/Haplotype.h:12-50,
 *  HapBlock.h:18-148, RepeatBlock.h:15-70) plus its pooled reads
 *  (read_pooler.h:14-53).  Everything is CSR-style: an offsets array of
 *  length count+1 indexes into a packed payload array.
 *
 *  Haplotype h of a locus selects one option per block by the reference's
 *  reflected mixed-radix Gray code with block 0 varying fastest
 *  (Haplotype.cpp:123-196); option 0 of a block is its reference sequence.
 * ------------------------------------------------------------------------- */
typedef struct hipstr_align_batch {
  int32_t n_loci;
  int32_t n_blocks;      /* total over loci */
  int32_t n_options;     /* total over blocks */
  int32_t n_pools;       /* total pooled reads over loci */
  int64_t n_haps;        /* total haplotypes over loci = sum prod(options per block) */

  /* per locus, length n_loci+1 */
  const int32_t* locus_block_off;   /* blocks  [off[l], off[l+1]) belong to locus l   */
  const int32_t* locus_pool_off;    /* pools   [off[l], off[l+1]) belong to locus l   */
  const int64_t* locus_hap_off;     /* haplotype-mask entries of locus l              */
  const int64_t* locus_out_off;     /* LL outputs of locus l: [P_l][H_l] row-major,
                                       pool-major / haplotype-minor like
                                       HapAligner.cpp:324 (aln_probs + i*num_combs)   */

  /* per block, length n_blocks (+1 for the offsets) */
  const int32_t* block_period;      /* 0 = flank block (HapBlock); >0 = repeat block
                                       (RepeatBlock) with this motif period          */
  const int32_t* block_opt_off;     /* options [off[b], off[b+1]) belong to block b   */
  const double*  block_stutter;     /* [n_blocks][6]: inframe_geom, inframe_up,
                                       inframe_down, outframe_geom, outframe_up,
                                       outframe_down (stutter_model.h:36-37); ignored
                                       for flank blocks                              */

  /* per option, length n_options+1 */
  const int32_t* opt_seq_off;       /* chars [off[o], off[o+1]) of opt_seq            */
  const char*    opt_seq;           /* allele sequences, forward strand orientation   */

  /* per pooled read, length n_pools (+1) */
  const int32_t* pool_seq_off;      /* chars [off[p], off[p+1]) of pool_bases/quals   */
  const char*    pool_bases;        /* read sequence (ASCII)                          */
  const char*    pool_quals;        /* Phred+33 base qualities (ASCII)                */
  const int32_t* pool_seed;         /* seed base index, or -1 = no seed: the read gets
                                       LL 0.0 for every haplotype (HapAligner.cpp:333) */

  /* optional masks (NULL = everything) */
  const uint8_t* realign_pool;      /* [n_pools]: 0 = leave this pool's outputs untouched
                                       (HapAligner.cpp:326-329)                       */
  const uint8_t* realign_hap;       /* [n_haps]: 0 = leave this column untouched and
                                       break DP-row reuse (HapAligner.cpp:615-619)    */
} hipstr_align_batch_t;

typedef struct hipstr_ctx hipstr_ctx_t;

/* --- context ------------------------------------------------------------- */

/* Create a context on CUDA device `device`.  Uploads the constant tables the
 * reference builds in precompute_integer_logs() (mathops.cpp:15-19),
 * init_alignment_model() (SeqAlignment/AlignmentModel.cpp:20-32) and
 * BaseQuality() (base_quality.h:29-38). */
hipstr_status_t hipstr_create(int device, hipstr_ctx_t** out_ctx);
void            hipstr_destroy(hipstr_ctx_t* ctx);
const char*     hipstr_last_error(const hipstr_ctx_t* ctx);
const char*     hipstr_version(void);

/* Use an externally owned CUDA stream (cudaStream_t passed as void*) for all
 * subsequent work of this context; NULL restores the context's own stream. */
hipstr_status_t hipstr_set_stream(hipstr_ctx_t* ctx, void* cuda_stream);

/* --- a5: seed selection (host integer logic) -----------------------------
 * Replaces HapAligner::calc_seed_base (SeqAlignment/HapAligner.cpp:270-318)
 * + calc_best_seed_position (:238-264) for n_reads alignments of ONE locus.
 * CIGAR ops of read r are cigar_type/cigar_len[cigar_off[r]..cigar_off[r+1]),
 * types are the reference's '=', 'X', 'I', 'D' (AlignmentData.h:12-27).
 * first_block_start / last_block_end and the repeat block [start,end) lists
 * are the genomic coordinates HapBlock::start()/end() return. */
hipstr_status_t hipstr_calc_seeds(int32_t n_reads, const int32_t* read_start,
                                  const int32_t* read_len,
                                  const int32_t* cigar_off, const char* cigar_type,
                                  const int32_t* cigar_len,
                                  int32_t first_block_start, int32_t last_block_end,
                                  int32_t n_repeats, const int32_t* repeat_start,
                                  const int32_t* repeat_end, int32_t* out_seed);

/* --- a2: read pooling (host byte logic) ------------------------------------
 * Replaces ReadPooler::add_alignment (read_pooler.cpp:3-20) + ReadPooler::pool
 * (read_pooler.h:42-48) + BaseQuality::median_base_qualities
 * (base_quality.cpp:11-28) for the n_reads reads of ONE locus.  Reads with an
 * identical base sequence share a pool; pools are numbered in order of first
 * appearance; a pool inherits the coordinates / CIGAR of its first member
 * (pool_first_read) and, per position, the UPPER median (sorted[n/2]) of its
 * members' quality bytes.
 *   seq_off [n_reads+1] offsets into bases/quals
 *   pool_index [n_reads] out; *n_pools out; pool_first_read [n_reads] capacity
 *   pool_seq_off [n_reads+1] capacity; pool_bases/pool_quals: capacity
 *   seq_off[n_reads] bytes each. */
hipstr_status_t hipstr_pool_reads(int32_t n_reads, const int32_t* seq_off, const char* bases,
                                  const char* quals, int32_t* pool_index, int32_t* n_pools,
                                  int32_t* pool_first_read, int32_t* pool_seq_off,
                                  char* pool_bases, char* pool_quals);

/* --- a4,a6-a10: read x haplotype HMM alignment (kernel K1) ----------------
 * Replaces HapAligner::process_reads (SeqAlignment/HapAligner.h:86-87, impl
 * HapAligner.cpp:320-343 -> process_read :573-709 -> align_seq_to_hap :26-161,
 * compute_aln_logprob :163-231, StutterAlignerClass.cpp:12-162) for every
 * pooled read of every locus of the batch in one call.
 *   ll_out       [locus_out_off[n_loci]] doubles; entry (l,p,h) at
 *                locus_out_off[l] + p*H_l + h.  Entries masked out by
 *                realign_pool / realign_hap are left untouched (so the caller's
 *                buffer content survives, as in the reference).
 *   seed_hap_pos optional (may be NULL), same indexing, int32: haplotype
 *                position the seed base aligns to in the best placement
 *                (max_index of compute_aln_logprob, HapAligner.cpp:184-222).
 * Host buffers in, host buffers out; copies are part of the call. */
hipstr_status_t hipstr_align_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                        double* ll_out, int32_t* seed_hap_pos);

/* Stage a batch in device memory (host -> HBM) and return a handle; then run
 * the kernels any number of times on the resident copy.  ll_dev is a DEVICE
 * pointer of locus_out_off[n_loci] doubles (seed_hap_pos_dev may be NULL). */
typedef struct hipstr_dev_batch hipstr_dev_batch_t;
hipstr_status_t hipstr_upload_batch(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                    hipstr_dev_batch_t** out_handle);
hipstr_status_t hipstr_align_batch_dev(hipstr_ctx_t* ctx, const hipstr_dev_batch_t* handle,
                                       double* ll_dev, int32_t* seed_hap_pos_dev);
void            hipstr_free_batch(hipstr_ctx_t* ctx, hipstr_dev_batch_t* handle);
/* Number of (pooled read, haplotype) alignments the batch performs (mask aware,
 * seedless reads excluded): the count of `*prob_ptr = LL` executions at
 * HapAligner.cpp:632. */
int64_t         hipstr_batch_num_alignments(const hipstr_align_batch_t* batch);
/* Kernel launches issued by the last align call of this context, and the device
 * time in ms of the last align_batch_dev call when timing is enabled
 * (CUDA events on the context's stream). */
int32_t         hipstr_last_launch_count(const hipstr_ctx_t* ctx);
hipstr_status_t hipstr_enable_timing(hipstr_ctx_t* ctx, int enable);
float           hipstr_last_kernel_ms(const hipstr_ctx_t* ctx);

/* --- a3 tail: pool -> read scatter and mate merge (kernel K2) -------------
 * Replaces the tail of SeqStutterGenotyper::calc_hap_aln_probs
 * (seq_stutter_genotyper.cpp:530-564) for one locus: copies pool LLs to member
 * reads for realigned haplotypes (copy_read / realign_hap may be NULL = all),
 * then replaces both mates' entries by their sum where second_mate[r] != 0.
 * read_ll is [n_reads][n_haps] row-major and is updated IN PLACE. */
hipstr_status_t hipstr_scatter_pool_lls_host(hipstr_ctx_t* ctx, int32_t n_reads, int32_t n_haps,
                                             const double* pool_ll, const int32_t* pool_seed,
                                             const int32_t* pool_index, const uint8_t* second_mate,
                                             const uint8_t* copy_read, const uint8_t* realign_hap,
                                             double* read_ll, int32_t* read_seed);

/* --- a13,a14: genotype posteriors (kernel K3) ------------------------------
 * Replaces Genotyper::calc_log_sample_posteriors (genotyper.h:72, impl
 * genotyper.cpp:44-80, priors :20-42) and get_optimal_haplotypes (:82-97) for a
 * batch of loci.  Reads are sample-major within a locus (genotyper.h:104-112).
 *   locus_read_off [n_loci+1], locus_sample_off [n_loci+1], n_haps [n_loci],
 *   haploid [n_loci] (0/1)
 *   read_ll    packed per locus [R_l][H_l]; log_p1/log_p2/sample_label/weight [R]
 *   post_out   packed per locus [S_l][H_l][H_l] normalised log posteriors
 *   sample_ll_out [S_total] per-sample total LL (sample_total_LLs_)
 *   best_out   [S_total][2] argmax diplotype (first maximum wins, strict >)
 *   returns per-locus total LL in total_ll_out [n_loci] (may be NULL). */
hipstr_status_t hipstr_posteriors_host(hipstr_ctx_t* ctx, int32_t n_loci,
                                       const int32_t* locus_read_off, const int32_t* locus_sample_off,
                                       const int32_t* n_haps, const uint8_t* haploid,
                                       const double* read_ll, const double* log_p1, const double* log_p2,
                                       const int32_t* sample_label, const int32_t* read_weight,
                                       double* post_out, double* sample_ll_out, int32_t* best_out,
                                       double* total_ll_out);

/* --- a1 first pass: align + scatter + posteriors for a batch of loci --------
 * One call = what SeqStutterGenotyper::genotype does at
 * seq_stutter_genotyper.cpp:638-639 (calc_hap_aln_probs(all haplotypes) ->
 * calc_log_sample_posteriors()) for MANY loci at once: K1 over every pooled
 * read x haplotype, K2 pool->read scatter + mate merge, K3 posteriors.
 *
 * hipstr_reads_batch_t is the read-level view of the same loci: the Genotyper /
 * SeqStutterGenotyper member arrays (genotyper.h:21-46,
 * seq_stutter_genotyper.h:34-68), packed over loci. */
typedef struct hipstr_reads_batch {
  const int32_t* locus_read_off;    /* [n_loci+1] reads of locus l: [off[l], off[l+1])        */
  const int32_t* locus_sample_off;  /* [n_loci+1]                                             */
  const int32_t* pool_index;        /* [R] pool_index_: pool of the read, local to its locus  */
  const int32_t* sample_label;      /* [R] sample_label_, local to its locus, non-decreasing  */
  const uint8_t* second_mate;       /* [R] second_mate_ (the read before it is its mate)      */
  const int32_t* read_weight;       /* [R] read_weights_ (0 for second mates)                 */
  const double*  log_p1;            /* [R] SNP-phasing log-likelihoods                        */
  const double*  log_p2;            /* [R]                                                    */
  const uint8_t* haploid;           /* [n_loci]                                               */
  const uint8_t* copy_read;         /* [R] or NULL = every read (calc_hap_aln_probs copy_read) */
} hipstr_reads_batch_t;

typedef struct hipstr_genotype_out {
  double*  read_ll;    /* [sum R_l*H_l] log_aln_probs_; entries the masks exclude keep their content */
  int32_t* read_seed;  /* [R] seed_positions_, may be NULL                                   */
  double*  post;       /* [sum S_l*H_l^2] log_sample_posteriors_                             */
  double*  sample_ll;  /* [S] sample_total_LLs_                                              */
  int32_t* best;       /* [S][2] get_optimal_haplotypes, may be NULL                         */
  double*  total_ll;   /* [n_loci] return value of calc_log_sample_posteriors, may be NULL   */
} hipstr_genotype_out_t;

/* HOST buffers in and out (end-to-end path: flatten, H2D, K1-K3, D2H). */
hipstr_status_t hipstr_genotype_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                           const hipstr_reads_batch_t* reads,
                                           const hipstr_genotype_out_t* out);
/* Resident variant: inputs staged once in HBM, `out` holds DEVICE pointers. */
typedef struct hipstr_dev_genotype hipstr_dev_genotype_t;
hipstr_status_t hipstr_upload_genotype_batch(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                             const hipstr_reads_batch_t* reads,
                                             hipstr_dev_genotype_t** out_handle);
hipstr_status_t hipstr_genotype_batch_dev(hipstr_ctx_t* ctx, const hipstr_dev_genotype_t* handle,
                                          const hipstr_genotype_out_t* dev_out);
void            hipstr_free_genotype_batch(hipstr_ctx_t* ctx, hipstr_dev_genotype_t* handle);

/* --- a14: genotypes and likelihoods from the posteriors (kernel K3b) ---------
 * Replaces Genotyper::extract_genotypes_and_likelihoods (genotyper.h:140-147, impl
 * genotyper.cpp:129-251, with calc_PLs :99-104 and calc_gl_diff :106-127) for a batch
 * of loci: haplotype posteriors are marginalised to STR-allele genotypes through
 * hap_to_allele (SeqStutterGenotyper::haps_to_alleles, seq_stutter_genotyper.cpp:219-227).
 *   n_variants  [n_loci]  number of alleles V_l of the STR block
 *   hap_to_allele packed per locus, [H_l] values in [0, V_l)
 *   post / sample_ll  outputs of hipstr_posteriors_host / hipstr_genotype_batch_*
 * Per sample outputs (all required unless noted):
 *   best_hap [S][2], best_gt [S][2]
 *   log_phased [S], log_unphased [S]            genotype posteriors (PQ / Q before exp)
 *   hap_log_phased [S], hap_log_unphased [S]
 *   gl   packed per locus [S_l][G_l], G_l = V(V+1)/2 (diploid, order (0,0),(1,0),(1,1),...) or V (haploid)
 *   phased_gl packed per locus [S_l][V_l*V_l] (diploid) or [S_l][V_l] (haploid)
 *   gl_diff [S]; pl packed like gl (int32). */
hipstr_status_t hipstr_extract_genotypes_host(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* locus_sample_off,
                                              const int32_t* n_haps, const int32_t* n_variants,
                                              const int32_t* hap_to_allele, const uint8_t* haploid,
                                              const double* post, const double* sample_ll,
                                              int32_t* best_hap, int32_t* best_gt, double* log_phased,
                                              double* log_unphased, double* hap_log_phased,
                                              double* hap_log_unphased, double* gl, double* phased_gl,
                                              double* gl_diff, int32_t* pl);

/* --- a16: alignment traceback (kernel K5) ------------------------------------
 * Replaces HapAligner::trace_optimal_aln (SeqAlignment/HapAligner.h:93, impl
 * HapAligner.cpp:711-722 -> process_read(retrace_aln = true) :636-690 -> retrace :363-571)
 * for a list of (pooled read, haplotype) pairs of a batch: the read is realigned to that ONE
 * haplotype keeping the full matrices, the best seed placement is found, and the most likely
 * path is walked back on both sides of the seed (ties within TRACE_LL_TOL = 0.001 resolved as
 * the reference does, :345-361).  What the reference stores as strings in AlignmentTrace
 * (SeqAlignment/AlignmentTraceback.h:10-115) comes back as index ranges into the read:
 *   str_seq(b) / flank_seq(b) == read[span_start[b] .. span_start[b] + span_len[b]).
 * Re-expressing the trace against the reference genome (stitch_alignment_trace,
 * AlignmentTraceback.cpp:55-144) needs the haplotype-vs-reference alignments of
 * Haplotype::aln_haps_to_ref and stays with the caller.
 *   block_start [n_blocks] genomic start of every haplotype block (HapBlock::start())
 *   trace_pool / trace_hap [n_traces]: global pool index, haplotype index local to its locus */
#define HIPSTR_MAX_BLOCKS_PER_LOCUS 8
#define HIPSTR_MAX_TRACE_INDELS 16
#define HIPSTR_MAX_TRACE_SNPS 32
#define HIPSTR_NO_STR_DATA (-2147483647 - 1)
typedef struct hipstr_trace_out {
  int32_t  aln_stride;    /* bytes per trace in hap_aln, >= read length + haplotype length + 1      */
  char*    hap_aln;       /* [n_traces][aln_stride] 'M','I','D','S' ops, NUL-terminated (hap_aln()) */
  int32_t* seed_hap_pos;  /* [n_traces] haplotype position of the seed base (max_index)            */
  int32_t* stutter_size;  /* [n_traces][8] per block: stutter_size(b), HIPSTR_NO_STR_DATA if none  */
  int32_t* span_start;    /* [n_traces][8] per block: first read base of str_seq(b) / flank_seq(b) */
  int32_t* span_len;      /* [n_traces][8] per block: its length (0 = empty)                       */
  int32_t* flank_ins;     /* [n_traces] flank_ins_size()                                           */
  int32_t* flank_del;     /* [n_traces] flank_del_size()                                           */
  int32_t* n_indels;      /* [n_traces] entries of flank_indel_data(), in the reference's order: the TRUE count */
  int32_t* indels;        /* [n_traces][16][2] (position, size): the first min(n_indels, 16) entries; a trace
                             with more gets its full lists from hipstr_trace_flank_lists                       */
  int32_t* n_snps;        /* [n_traces] entries of flank_snp_data(): the TRUE count                */
  int32_t* snps;          /* [n_traces][32][2] (position, read base character): the first min(n_snps, 32)      */
} hipstr_trace_out_t;

hipstr_status_t hipstr_trace_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                        const int32_t* block_start, int32_t n_traces,
                                        const int32_t* trace_pool, const int32_t* trace_hap,
                                        const hipstr_trace_out_t* out);

/* The COMPLETE flank indel / flank SNP lists of one trace (AlignmentTrace::flank_indel_data / flank_snp_data,
 * SeqAlignment/AlignmentTraceback.h:29-33).  hipstr_trace_batch_host returns these lists in fixed-size slots
 * (HIPSTR_MAX_TRACE_INDELS / HIPSTR_MAX_TRACE_SNPS entries per trace) next to the TRUE counts n_indels / n_snps; when
 * a count exceeds its slot -- a chimeric or mismapped read -- only the first entries were stored, and the caller gets the
 * whole lists here, rebuilt on the host from the trace's operation string (the accounting half of HapAligner::retrace,
 * HapAligner.cpp:363-571, driven by the operations K5 chose), identical to the reference's.
 *   batch / block_start / pool / hap   as passed to hipstr_trace_batch_host (hap local to the pool's locus)
 *   hap_aln, seed_hap_pos, stutter_size [8]   that trace's outputs
 *   own_quals   NULL = the pool's qualities; else the qualities the trace was computed with (a read traced with its own)
 *   indels [cap_indels][2] (position, size), snps [cap_snps][2] (position, base character); *n_* = the true counts */
hipstr_status_t hipstr_trace_flank_lists(const hipstr_align_batch_t* batch, const int32_t* block_start, int32_t pool, int32_t hap,
                                         const char* hap_aln, int32_t seed_hap_pos, const int32_t* stutter_size,
                                         const char* own_quals, int32_t cap_indels, int32_t* n_indels, int32_t* indels,
                                         int32_t cap_snps, int32_t* n_snps, int32_t* snps);

/* --- a16 (host part): trace -> alignment against the reference genome ---------
 * Replaces stitch_alignment_trace + stitch (SeqAlignment/AlignmentTraceback.cpp:5-144), the
 * last step of process_read(retrace_aln = true) (HapAligner.cpp:686-688): the read-vs-haplotype
 * operation string of K5 (hap_aln, seed_hap_pos) is composed with the haplotype-vs-reference
 * alignment string of Haplotype::aln_haps_to_ref ('M','I','D' per column, Haplotype.cpp:58-86) into
 * the read's alignment against the reference: start/stop coordinates, CIGAR ('M','I','D','S'
 * runs) and the gapped read string ('-' for deleted reference bases, soft-clipped bases dropped).
 *   hap_start   genomic start of the haplotype's first block (HapBlock::start())
 *   cigar_cap   capacity of cigar_type / cigar_len; *n_cigar receives the number of runs
 *   aln_cap     capacity of alignment (NUL-terminated on return)
 * cigar_type / cigar_len / alignment may all be NULL when only start / stop are wanted.
 * Pure host string logic; HIPSTR_ERR_BAD_ARG on inconsistent inputs (the reference dies). */
hipstr_status_t hipstr_stitch_trace(int32_t hap_start, const char* hap_aln_to_ref, const char* read_aln_to_hap,
                                    int32_t seed_hap_pos, int32_t seed_base, const char* read_bases,
                                    int32_t* start, int32_t* stop, int32_t cigar_cap, char* cigar_type,
                                    int32_t* cigar_len, int32_t* n_cigar, int32_t aln_cap, char* alignment);

/* The span of hipstr_stitch_trace (start / stop only) for MANY traces against the same haplotype: the haplotype's operation
 * string is indexed once (hipstr_hap_aln_index: three int32 tables of hap_aln_len + 1 entries each -- non-'D' columns before a
 * column, non-'I' columns before a column, column of the k-th non-'D' operation), and a trace then jumps over the part of the
 * walk that consumes haplotype bases instead of stepping through it; only the operations of the read beyond its outermost
 * aligned base on either side are stepped through.  Same results and status as hipstr_stitch_trace with NULL string buffers
 * (tests/test_trace.py compares them on random operation strings).  index must hold 3 * (hap_aln_len + 1) entries. */
hipstr_status_t hipstr_hap_aln_index(const char* hap_aln_to_ref, int32_t hap_aln_len, int32_t* index);
hipstr_status_t hipstr_trace_span(int32_t hap_start, const char* hap_aln_to_ref, int32_t hap_aln_len, const int32_t* index,
                                  const char* read_aln_to_hap, int32_t seed_hap_pos, int32_t seed_base, int32_t* start,
                                  int32_t* stop);

/* --- a17 / seam B4: EM stutter-model learner (kernel K4) --------------------
 * Replaces EMStutterGenotyper(...) + train(...) + get_stutter_model()
 * (em_stutter_genotyper.h:50-117, em_stutter_genotyper.cpp:10-226) for a batch of
 * loci, each an independent EM problem over read STR lengths.  Reads are
 * sample-major per locus like everywhere else.
 *   num_bps      [R] base-pair size of each read's STR (the ctor's num_bps)
 *   motif_len    [n_loci], ref_allele [n_loci] (bp size of the reference allele)
 * Outputs per locus: params [6] = inframe_geom, inframe_up, inframe_down,
 * outframe_geom, outframe_up, outframe_down (StutterModel ctor order,
 * stutter_model.h:36-37); converged = train()'s return value; iterations run;
 * final total log-likelihood. */
typedef struct hipstr_em_batch {
  int32_t n_loci;
  const int32_t* locus_read_off;    /* [n_loci+1] */
  const int32_t* locus_sample_off;  /* [n_loci+1] */
  const int32_t* num_bps;           /* [R] */
  const int32_t* sample_label;      /* [R] local to the locus, non-decreasing */
  const double*  log_p1;            /* [R] */
  const double*  log_p2;            /* [R] */
  const int32_t* motif_len;         /* [n_loci] */
  const int32_t* ref_allele;        /* [n_loci] */
  const uint8_t* haploid;           /* [n_loci] */
} hipstr_em_batch_t;

hipstr_status_t hipstr_em_train_host(hipstr_ctx_t* ctx, const hipstr_em_batch_t* batch, int32_t max_iter,
                                     double min_LL_abs_change, double min_LL_frac_change,
                                     double* params_out, uint8_t* converged_out, int32_t* iters_out,
                                     double* ll_out);

/* --- a1, a15 / seam B1: the per-locus genotyping loop, many loci at once --------
 * Replaces SeqStutterGenotyper(...) + genotype() (seq_stutter_genotyper.h:143-189; init()
 * .cpp:486-517, genotype() :603-671, id_and_align_to_stutter_alleles :570-601,
 * get_stutter_candidate_alleles :843-879, get_unused_alleles :229-315, add_and_remove_alleles
 * :324-415, retrace_alignments :805-841) for a BATCH of loci.  The reference builds one object
 * per locus and loops align -> trace -> add / remove alleles serially; here all loci advance
 * in lockstep rounds and every round is at most three device calls (hipstr_trace_batch_host,
 * hipstr_genotype_batch_host masked to the new haplotypes, hipstr_posteriors_host), with the
 * reference's per-locus decisions made on the host in between.  C++ callers use
 * hipstr::GenotyperBatch / hipstr::SeqStutterGenotyper (hipstr_b200/host/seq_stutter_genotyper.h).
 *
 * Inputs are what seam B1 receives after haplotype generation: the haplotype blocks of every
 * locus (only the locus/block/option fields of hipstr_align_batch_t are read) with their
 * reference coordinates, and the un-pooled, left-aligned reads, sample-major per locus
 * (std::vector<Alignment>, genotyper.h:104-112).  Pooling, second-mate detection (adjacent reads
 * with equal name_id, .cpp:499) and seeding happen inside.  ref_vcf == NULL semantics (no
 * reference panel); flank re-assembly (assemble_flanks .cpp:40-217 over DebruijnGraph,
 * debruijn_graph.cpp, directed_graph.cpp) runs on the host between rounds when requested.
 * A locus the reference would skip (too many haplotypes, repetitive flanks) ends with
 * locus_ok = 0 and its reason in the locus log; the call itself still returns HIPSTR_OK. */
typedef struct hipstr_locus_reads {
  const int32_t* locus_read_off;    /* [n_loci+1]                                              */
  const int32_t* locus_sample_off;  /* [n_loci+1]                                              */
  const int32_t* read_seq_off;      /* [R+1] offsets into bases / quals                        */
  const char*    bases;             /* Alignment::get_sequence()                               */
  const char*    quals;             /* Alignment::get_base_qualities() (Phred+33)              */
  const int32_t* read_start;        /* [R] Alignment::get_start()                              */
  const int32_t* cigar_off;         /* [R+1]                                                   */
  const char*    cigar_type;        /* '=', 'X', 'I', 'D' (AlignmentData.h:12-27)              */
  const int32_t* cigar_len;
  const int32_t* sample_label;      /* [R] local to the locus, non-decreasing                  */
  const int32_t* name_id;           /* [R] stand-in for the read name: equal ids on adjacent reads = mates */
  const double*  log_p1;            /* [R]                                                     */
  const double*  log_p2;            /* [R]                                                     */
  const uint8_t* haploid;           /* [n_loci]                                                */
  const uint8_t* rev_strand;        /* [R] Alignment::is_from_reverse_strand(), NULL = all forward */
  const int32_t* read_stop;         /* [R] Alignment::get_stop(): HipSTR keeps the LAST aligned reference position
                                       (inclusive; AlignmentOps.cpp:56-57,106).  NULL = derived from the CIGAR as
                                       read_start + reference bases consumed - 1                              */
  const uint8_t* use_for_haps;      /* [R] Alignment::use_for_hap_generation(0) (AlignmentData.h:116-128; set from the
                                       "PF" tag of the read filters, genotyper_bam_processor.cpp:91-94): only these
                                       reads propose candidate alleles in build_haplotype (seq_stutter_genotyper.cpp:
                                       438-442); every read is still aligned and genotyped.  NULL = all reads   */
} hipstr_locus_reads_t;

typedef struct hipstr_genotyper hipstr_genotyper_t;
/* ctx may be NULL for the two constructors (pooling, seeding and haplotype generation are host
 * work and can be inspected without a GPU); genotype() / write_vcf then return
 * HIPSTR_ERR_NO_DEVICE -- there is no CPU alignment path. */
hipstr_status_t hipstr_genotyper_create(hipstr_ctx_t* ctx, const hipstr_align_batch_t* blocks,
                                        const int32_t* block_start, const int32_t* block_end,
                                        const hipstr_locus_reads_t* reads, hipstr_genotyper_t** out);
/* The full constructor of seam B1: the haplotype blocks are generated from the reads themselves
 * (SeqStutterGenotyper::build_haplotype .cpp:422-484 over HaplotypeGenerator::add_haplotype_block /
 * fuse_haplotype_blocks, SeqAlignment/HaplotypeGenerator.cpp:12-366), one STR region per locus:
 *   region_start / region_stop / period [n_loci]  the Region;  chrom_seq [n_loci] its chromosome
 *   stutter [n_loci][6]  the locus' StutterModel parameters (block_stutter order)
 * A locus whose haplotype construction fails ("No spanning alignments", too near the chromosome
 * end) is kept uninitialised: genotype() reports locus_ok = 0 for it, like the reference. */
hipstr_status_t hipstr_genotyper_create_from_reads(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* region_start,
                                                   const int32_t* region_stop, const int32_t* period,
                                                   const char* const* chrom_seq, const double* stutter,
                                                   const hipstr_locus_reads_t* reads, hipstr_genotyper_t** out);
/* The same constructor with ref_vcf != NULL (the --ref-vcf reference panel, seq_stutter_genotyper.cpp:445-459): the
 * alleles of locus l are alleles[allele_off[l] .. allele_off[l+1]) starting at allele_pos[l] (0-based; what
 * read_vcf_alleles returns, src/vcf_input.cpp:21-50; allele 0 must equal the chromosome there), entered through
 * HaplotypeGenerator::add_vcf_haplotype_block (HaplotypeGenerator.cpp:256-284).  allele_pos[l] < 0 = the record could not
 * be read: the locus fails like in the reference.  genotype() then keeps the allele set fixed -- no stutter-allele
 * discovery and no pruning (:641-665, :204); flank assembly still runs. */
hipstr_status_t hipstr_genotyper_create_with_ref_alleles(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* region_start,
                                                         const int32_t* region_stop, const int32_t* period,
                                                         const char* const* chrom_seq, const double* stutter,
                                                         const hipstr_locus_reads_t* reads, const int32_t* allele_pos,
                                                         const int32_t* allele_off, const char* const* alleles,
                                                         hipstr_genotyper_t** out);
void            hipstr_genotyper_destroy(hipstr_genotyper_t* g);
const char*     hipstr_genotyper_last_error(const hipstr_genotyper_t* g);
/* genotype(max_total_haplotypes, max_flank_haplotypes, min_flank_freq) of every locus
 * (seq_stutter_genotyper.h:189; defaults 1000 / 4 / 0.01, genotyper_bam_processor.h:110-112);
 * reassemble_flanks is the constructor's flag (the reference always passes true,
 * genotyper_bam_processor.cpp:229).  locus_ok [n_loci] = genotype()'s return value. */
hipstr_status_t hipstr_genotyper_genotype(hipstr_genotyper_t* g, int32_t max_total_haplotypes,
                                          int32_t max_flank_haplotypes, double min_flank_freq,
                                          int32_t reassemble_flanks, uint8_t* locus_ok);
/* recompute_stutter_models(logger, max_total_haplotypes, max_flank_haplotypes, min_flank_freq, max_em_iter,
 * abs_ll_converge, frac_ll_converge) (seq_stutter_genotyper.h:195-196, impl .cpp:1541-1583; defaults 100 / 0.01 /
 * 0.001, genotyper_bam_processor.h:106-108) for every locus whose genotype() succeeded: the STR sizes of the
 * maximum-likelihood alignments train a new stutter model per repeat block (ONE hipstr_em_train_host call for all
 * loci), then genotype() runs again.  locus_ok = the method's return value per locus. */
hipstr_status_t hipstr_genotyper_recompute_stutter_models(hipstr_genotyper_t* g, int32_t max_total_haplotypes,
                                                          int32_t max_flank_haplotypes, double min_flank_freq,
                                                          int32_t max_em_iter, double abs_ll_converge,
                                                          double frac_ll_converge, uint8_t* locus_ok);
/* alignments (pooled read x haplotype) and traces computed so far, lockstep rounds run */
hipstr_status_t hipstr_genotyper_stats(const hipstr_genotyper_t* g, int64_t* n_alignments, int64_t* n_traces,
                                       int32_t* n_rounds);
/* wall-clock seconds by stage, seconds9 = {construction, per-locus host decisions, trace device calls,
 * trace stitching, alignment calls (K1+K2+K3 with packing), posterior calls, VCF formatting, and of the
 * alignment calls: packing the loci into one batch, unpacking the results} */
hipstr_status_t hipstr_genotyper_timing(const hipstr_genotyper_t* g, double* seconds9);
/* the host-decision seconds split by phase (summed over loci): {align-all set-up, stutter-allele discovery, uncalled
 * pruning, unspanned pruning, flank assembly, post-assembly pruning, done, failed} */
hipstr_status_t hipstr_genotyper_phase_timing(const hipstr_genotyper_t* g, double* seconds8);
/* info[8] = {blocks, haplotypes (num_alleles_), reads, samples, pools, total options,
 *            total allele bytes, alignment rounds of this locus} */
hipstr_status_t hipstr_genotyper_locus_info(const hipstr_genotyper_t* g, int32_t locus, int32_t* info);
/* current haplotype blocks: options per block, CSR offsets [total options + 1] and bytes */
hipstr_status_t hipstr_genotyper_locus_blocks(const hipstr_genotyper_t* g, int32_t locus, int32_t* block_n_opts,
                                              int32_t* opt_seq_off, char* opt_seq);
/* member arrays: log_aln_probs_ [R*H], seed_positions_ [R], pool_index_ [R], log_sample_posteriors_
 * [S*H*H], sample_total_LLs_ [S], get_optimal_haplotypes [S*2], call_sample_[s].empty() [S]; any may be NULL */
hipstr_status_t hipstr_genotyper_locus_results(const hipstr_genotyper_t* g, int32_t locus, double* read_ll,
                                               int32_t* read_seed, int32_t* pool_index, double* post,
                                               double* sample_ll, int32_t* best, uint8_t* call_sample_ok);
int32_t         hipstr_genotyper_locus_log(const hipstr_genotyper_t* g, int32_t locus, char* out, int32_t cap);

/* --- a18: write_vcf_record for every genotyped locus -----------------------------
 * Replaces SeqStutterGenotyper::write_vcf_record (seq_stutter_genotyper.h:179-181, impl
 * seq_stutter_genotyper.cpp:984-1510) with get_alleles (:691-769), reorder_alleles (:673-689),
 * compute_allele_bias (:965-982), ExtractCigar (extract_indels.cpp:18-90) and
 * Genotyper::condense_read_counts (genotyper.h:51-64).  One K3b call marginalises the
 * posteriors of all loci, one K5 call traces the reads whose strand-assigned haplotype has no
 * cached trace, then the text is formatted on the host exactly as the reference prints it
 * (fixed, 2 decimals).  Loci whose genotype() failed produce no record (the reference skips
 * write_vcf_record for them, genotyper_bam_processor.cpp:232-246).  One STR region per locus.
 * The OUTPUT_* switches are the reference's static Genotyper flags (genotyper.cpp:336-343);
 * hipstr_vcf_default_options() returns their defaults.  HTML visualisation is not produced. */
typedef struct hipstr_vcf_loci {
  const char* const* chrom;          /* [n_loci] Region::chrom()                                */
  const char* const* name;           /* [n_loci] Region::name(), "" or NULL -> "."              */
  const int32_t* region_start;       /* [n_loci] Region::start()                                */
  const int32_t* region_stop;        /* [n_loci] Region::stop()                                 */
  const int32_t* period;             /* [n_loci] Region::period()                               */
  const char* const* chrom_seq;      /* [n_loci] chromosome sequence the coordinates index into */
  const char* const* locus_sample_names; /* [total samples] the genotyper's sample_names_, per locus */
  int32_t n_out_samples;             /* samples_to_genotype: the VCF's sample columns           */
  const char* const* out_sample_names;
} hipstr_vcf_loci_t;
typedef struct hipstr_vcf_options {
  int32_t output_gls, output_pls, output_phased_gls, output_allreads, output_mallreads, output_filters,
          output_haplotype_data;
  double max_flank_indel_frac;
} hipstr_vcf_options_t;
void hipstr_vcf_default_options(hipstr_vcf_options_t* o);
hipstr_status_t hipstr_genotyper_write_vcf(hipstr_genotyper_t* g, const hipstr_vcf_loci_t* loci,
                                           const hipstr_vcf_options_t* options);
/* the record of one locus (empty when the locus produced none); returns its length or -needed */
int32_t         hipstr_genotyper_locus_record(const hipstr_genotyper_t* g, int32_t locus, int32_t* pos, char* out, int32_t cap);

/* --- the locus list of seam B1 on every GPU of the box, from C++ ------------------
 * Replaces the region loop of BamProcessor::process_regions (src/bam_processor.cpp:550-617) around
 * GenotyperBamProcessor::analyze_reads_and_phasing (src/genotyper_bam_processor.cpp:229-246: constructor, genotype(),
 * write_vcf_record per locus) for a whole list of loci.  The list is cut into windows of `window_loci` consecutive
 * loci; one worker thread per (device, pipeline) owns a context on its device and pulls the next window when it has
 * finished one, heaviest windows (most reads) first, so no static split has to guess the cost of a locus.  Several
 * pipelines per device overlap the host stages of one window with the device stages of another.  There is no exchange
 * between devices; the records are kept per locus and read back in locus order, the order VCFWriter::add_vcf_record
 * needs (src/vcf_writer.h:33-35).
 *   devices [n_devices]       CUDA device ordinals; pipelines_per_device 1..8 (2-3 keep one GPU busy)
 *   next_window / user        NULL = the handle's own counter.  Otherwise the dealer: returns the next position in the
 *                             dealing order (0, 1, 2, ...; anything outside [0, windows) stops the worker) -- lets the
 *                             workers of several PROCESSES (one per GPU under torchrun) share one locus list through a
 *                             counter of their own (hipstr_multi_window_order is a pure function of the inputs).
 *   locus_ok [n_loci]         genotype()'s return value per locus (0 for windows another process took); may be NULL
 * Arguments otherwise as hipstr_genotyper_create_from_reads / _genotype (reassemble_flanks = 1) / _write_vcf. */
typedef struct hipstr_multi hipstr_multi_t;
typedef struct hipstr_vcf_writer hipstr_vcf_writer_t;   /* seam B5, declared below */
typedef int32_t (*hipstr_next_window_fn)(void* user);
hipstr_status_t hipstr_multi_create(int32_t n_devices, const int32_t* devices, int32_t pipelines_per_device, hipstr_multi_t** out);
void            hipstr_multi_destroy(hipstr_multi_t* m);
const char*     hipstr_multi_last_error(const hipstr_multi_t* m);
int32_t         hipstr_multi_num_workers(const hipstr_multi_t* m);
int32_t         hipstr_multi_num_windows(int32_t n_loci, int32_t window_loci);
/* order [windows]: the window dealt at position k (heaviest first; cost = reads of the window) */
hipstr_status_t hipstr_multi_window_order(int32_t n_loci, int32_t window_loci, const int32_t* locus_read_off, int32_t* order);
hipstr_status_t hipstr_multi_genotype(hipstr_multi_t* m, int32_t n_loci, const int32_t* region_start, const int32_t* region_stop,
                                      const int32_t* period, const char* const* chrom_seq, const double* stutter,
                                      const hipstr_locus_reads_t* reads, const hipstr_vcf_loci_t* vcf_loci,
                                      const hipstr_vcf_options_t* vcf_options, int32_t max_total_haplotypes,
                                      int32_t max_flank_haplotypes, double min_flank_freq, int32_t window_loci,
                                      hipstr_next_window_fn next_window, void* next_window_user, uint8_t* locus_ok);
/* the record of one locus (empty when the locus produced none or another process took its window); returns its length or -needed */
int32_t         hipstr_multi_locus_record(const hipstr_multi_t* m, int32_t locus, int32_t* pos, char* out, int32_t cap);
/* feed the records to a writer in locus order (VCFWriter::add_vcf_record) */
hipstr_status_t hipstr_multi_emit_records(const hipstr_multi_t* m, const hipstr_vcf_loci_t* loci, hipstr_vcf_writer_t* w);
/* alignments / traces computed, stage seconds summed over windows (layout of hipstr_genotyper_timing), and per worker
 * (device-major, pipeline-minor) the windows it processed and the seconds it was busy; any pointer may be NULL */
hipstr_status_t hipstr_multi_stats(const hipstr_multi_t* m, int64_t* n_alignments, int64_t* n_traces, double* seconds9,
                                   int32_t* windows_per_worker, double* busy_seconds_per_worker);
/* host->device and device->host bytes and kernel launches of the last hipstr_multi_genotype call, summed over every device
 * call of every window (each window's reads, haplotypes and results cross the bus per round: see DESIGN.md 7) */
hipstr_status_t hipstr_multi_traffic(const hipstr_multi_t* m, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* gpu_launches);

/* Pure host arithmetic of write_vcf_record, exported so that it can be checked on its own:
 *   hipstr_allele_bias       compute_allele_bias (seq_stutter_genotyper.cpp:965-982): log10 of the two-sided
 *                            binomial p-value of the read split (the reference uses cephes bdtr, lib/cephes/bdtr.c)
 *   hipstr_fisher_two_sided  the `two` output of kt_fisher_exact (htslib 1.9 kfunc.c:196-279)
 *   hipstr_extract_cigar     ExtractCigar (extract_indels.cpp:18-90): 1 and *bp_diff if the read spans the window */
double  hipstr_allele_bias(int32_t hap_a_reads, int32_t hap_b_reads);
double  hipstr_fisher_two_sided(int32_t n11, int32_t n12, int32_t n21, int32_t n22);
int32_t hipstr_extract_cigar(const char* cigar_type, const int32_t* cigar_len, int32_t n, int32_t cigar_start,
                             int32_t region_start, int32_t region_end, int32_t* bp_diff);

/* The per-sample step of assemble_flanks (seq_stutter_genotyper.cpp:49-97) on its own: De Bruijn graph of the reference
 * flank (weight 2) and the sample's flank sequences for the smallest k in [min_kmer, max_kmer] that leaves the
 * reference acyclic (DebruijnGraph::calc_kmer_length) and the pruned graph acyclic with a clean source and sink, then
 * the best-first enumeration of source-to-sink paths (debruijn_graph.cpp:151-199, min weight 2).  Returns the number of
 * paths (sequence i at paths + i * path_cap, its bottleneck weight in weights[i]; *k_used = the k), -1 if the
 * reference flank is too repetitive, -3 if every k is cyclic, -2 on bad arguments.  Pure host logic. */
int32_t hipstr_flank_assemble(const char* ref_seq, int32_t n_seqs, const char* const* seqs, int32_t min_kmer,
                              int32_t max_kmer, int32_t* k_used, int32_t max_paths, int32_t path_cap, char* paths,
                              int32_t* weights);

/* Haplotype::aln_haps_to_ref for one haplotype (SeqAlignment/Haplotype.cpp:8-86 on top of
 * NeedlemanWunsch::Align, NeedlemanWunsch.cpp:84-423): one of 'M','I','D' per alignment column
 * of alt_hap against ref_hap -- the string hipstr_stitch_trace consumes.  Pure host logic. */
hipstr_status_t hipstr_hap_aln_to_ref(const char* ref_hap, const char* alt_hap, int32_t first_block_start,
                                      int32_t repeat_block_start, int32_t cap, char* out);

/* --- the arithmetic of read left-alignment (kernel K6) ------
 * Replaces NeedlemanWunsch::Align (SeqAlignment/NeedlemanWunsch.h:16-18, impl NeedlemanWunsch.cpp:384-423:
 * initMatrices :339-381, nw_helper :193-241, findOptimalStop / findOptimalStopEndPenalty :149-191,
 * traceAlignment :243-337) for a BATCH of (reference window, read) pairs -- what realign()
 * (SeqAlignment/AlignmentOps.cpp:14-100) runs once per distinct read sequence inside left_align_reads
 * (genotyper_bam_processor.cpp:38-102), and Haplotype::aln_haps_to_ref once per haplotype.
 *   ref_off / read_off [n_pairs+1] offsets into ref_seqs / read_seqs (A, C, G, T, anything else scores like N)
 *   use_ref_end_penalty  0 = the read may start and end anywhere in the window (realign), 1 = end-to-end
 *   ops  [n_pairs][ops_stride], ops_stride > longest window + longest read: one character per alignment column,
 *        'M' base against base, 'D' reference base against a gap (also the unaligned window ends),
 *        'I' read base against a gap; NUL-terminated.  ref_al / read_al / the CIGAR of the reference follow
 *        by walking the two sequences along it (hipstr_realign_read).
 *   ops_len [n_pairs] number of columns; score [n_pairs] the alignment score (float, like the reference)
 * Tie decisions are the reference's (bestIndex :125-147), so the strings are identical, not merely optimal. */
hipstr_status_t hipstr_nw_align_batch_host(hipstr_ctx_t* ctx, int32_t n_pairs, const int32_t* ref_off,
                                           const char* ref_seqs, const int32_t* read_off, const char* read_seqs,
                                           int32_t use_ref_end_penalty, int32_t ops_stride, char* ops,
                                           int32_t* ops_len, float* score);

/* GenotyperBamProcessor::left_align_reads (genotyper_bam_processor.cpp:38-102) for a batch of loci: every read is
 * trimmed to [trim_start, trim_stop] of its locus (BamAlignment::TrimAlignment, bam_io.cpp:384-477; the reference
 * passes region_group.start() - 40 (or 1) and region_group.stop() + 40; NULL = no trimming), reads whose CIGAR holds
 * only M / = are re-expressed with = / X (convertAlignment, AlignmentOps.cpp:102-167), the others are re-aligned
 * against chromosome[Position - 76, EndPosition + 74] once per distinct sequence of the locus (realign, :14-100) --
 * ALL Needleman-Wunsch alignments of the call in one K6 launch (a second one only for sequences whose first
 * alignment came back clipped) -- and later reads with a known sequence reuse that alignment (:62-79).
 *   raw  BAM-level alignments, sample-major per locus, in hipstr_locus_reads_t: read_start = Position(),
 *        read_stop = GetEndPosition() (BAM's exclusive end, REQUIRED), CIGAR operations M = X I D S H.
 * The result owns a hipstr_locus_reads_t of the left-aligned reads (HipSTR's conventions: = / X / I / D CIGARs,
 * inclusive stop), ready for hipstr_genotyper_create_from_reads; reads that are empty after trimming or fail to
 * realign are dropped (hipstr_left_aligned_source maps every kept read to its input index).  Region filters
 * (set_hap_gen_info, :90-92) stay with the caller. */
typedef struct hipstr_left_aligned hipstr_left_aligned_t;
hipstr_status_t hipstr_left_align_reads_host(hipstr_ctx_t* ctx, int32_t n_loci, const hipstr_locus_reads_t* raw,
                                             const char* const* chrom_seq, const int32_t* trim_start,
                                             const int32_t* trim_stop, hipstr_left_aligned_t** out);
const hipstr_locus_reads_t* hipstr_left_aligned_reads(const hipstr_left_aligned_t* h);
const int32_t* hipstr_left_aligned_source(const hipstr_left_aligned_t* h, int64_t* n_reads);
void hipstr_left_aligned_counts(const hipstr_left_aligned_t* h, int64_t* failed, int64_t* nw_alignments);
void hipstr_left_aligned_free(hipstr_left_aligned_t* h);
/* The host steps for ONE read, exported so that they can be checked without a GPU: returns -1 nothing left after
 * trimming, 1 converted, 3 "needs a Needleman-Wunsch alignment" (window = {start, length} of the chromosome window,
 * out_seq = the trimmed read; call again with nw_ops = the operation string of that alignment), 2 realigned, 0 failed. */
int32_t hipstr_left_align_one(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar,
                              const char* cigar_type, const int32_t* cigar_len, const char* chrom_seq, int32_t do_trim,
                              int32_t trim_start, int32_t trim_stop, const char* nw_ops, int32_t* window,
                              int32_t* out_pos, char* out_seq, char* out_qual, int32_t* n_out_cigar, char* out_ctype,
                              int32_t* out_clen);

/* --- seam B5: ordered VCF output (host) ---------------------------------------
 * Replaces VCFWriter::open / write_header / add_vcf_record / close (vcf_writer.h:63-82,
 * vcf_writer.cpp:7-36): records of a chromosome may arrive up to 50 bp out of order and are
 * held in a min-heap keyed by POS until no later record can precede them; chromosomes must
 * arrive grouped.  A path ending in ".gz" is written as BGZF (what the reference's
 * bgzfostream produces), anything else as plain text.  The C++ class with the reference's
 * method names is hipstr::VCFWriter (hipstr_b200/host/vcf_writer.h). */
/* Genotyper::get_vcf_header (src/genotyper.cpp:253-331): the header text of the STR VCF for these options -- file format,
 * ##command, ##reference, one ##contig line per FASTA sequence (FastaReader::write_all_contigs_to_vcf), the INFO / FORMAT
 * dictionary and the #CHROM line with the sample columns.  Returns the text length, or -(needed capacity). */
int64_t hipstr_vcf_header(const char* reference_path, const char* full_command, int32_t n_contigs, const char* const* contig_names,
                          const int64_t* contig_lengths, int32_t n_samples, const char* const* sample_names,
                          const hipstr_vcf_options_t* options, int64_t cap, char* out_text);
typedef struct hipstr_vcf_writer hipstr_vcf_writer_t;
/* feed the records of hipstr_genotyper_write_vcf to a writer, in locus order (VCFWriter::add_vcf_record) */
hipstr_status_t hipstr_genotyper_emit_records(const hipstr_genotyper_t* g, const hipstr_vcf_loci_t* loci,
                                              hipstr_vcf_writer_t* w);
hipstr_vcf_writer_t* hipstr_vcf_writer_open(const char* path);   /* NULL if the file cannot be created */
hipstr_status_t hipstr_vcf_writer_header(hipstr_vcf_writer_t* w, const char* header_text);
hipstr_status_t hipstr_vcf_writer_add_record(hipstr_vcf_writer_t* w, const char* chrom, int32_t pos,
                                             const char* record_text);
void            hipstr_vcf_writer_close(hipstr_vcf_writer_t* w);   /* flushes, writes the BGZF EOF block, frees */
/* the same with the outcome: HIPSTR_ERR_BAD_ARG if any write, the compression or the final fclose failed (full disk,
 * closed pipe) -- the file on disk is then incomplete; _header / _add_record report a failed write the same way */
hipstr_status_t hipstr_vcf_writer_finish(hipstr_vcf_writer_t* w);

/* --- section 8(f) row 4, first slice: SNP phasing log-likelihoods (K7) ----------
 * Replaces calc_het_snp_factors (src/snp_phasing_quality.cpp:92-120) -> add_log_phasing_probs (:65-90) ->
 * extract_bases_and_qualities (:4-63) with SNPTree::findContained (src/snp_tree.h:114-126), as called by
 * SNPBamProcessor::process_reads (src/snp_bam_processor.cpp:60-76), for all reads of a batch of loci in ONE
 * launch.  These are the log_p1 / log_p2 the genotyper takes (hipstr_locus_reads_t.log_p1 / log_p2).
 *
 * An ENTRY is one STR read, optionally followed by its mate (the paired overload): the terms of the read's
 * SNPs are added first, then the mate's, into the same two doubles, like the reference; results are
 * bit-identical.  A SNP SET is the position-sorted list of one sample's phased heterozygous SNPs (what
 * create_snp_trees puts into that sample's SNPTree); entry_snp_set = -1 for a sample without SNP
 * information leaves both values 0 (snp_bam_processor.cpp:77-84).  Alignments carry the fields of the
 * (possibly trimmed) BamAlignment: Position(), GetEndPosition() (exclusive), QueryBases(), Qualities(),
 * CigarData() with the BAM operation characters M = X D I S H.
 * counts [n_entries][4]: bases matching haplotype one, haplotype two, neither (the reference's
 * p1_match_count / p2_match_count / mismatch_count, summed by the caller), and a status that is 0 unless
 * the reference would have died on the entry (1 invalid CIGAR character, 2 CIGAR longer than the read,
 * 3 CIGAR shorter than the alignment span); any non-zero status makes the call return HIPSTR_ERR_BAD_ARG. */
typedef struct {
  int32_t n_entries;
  const int32_t* entry_aln_off;   /* [n_entries+1] alignments of entry e: [off[e], off[e+1]) */
  const int32_t* entry_snp_set;   /* [n_entries] index of the sample's SNP set, or -1 */
  int32_t n_alns;
  const int32_t* aln_pos;         /* [n_alns] */
  const int32_t* aln_end;         /* [n_alns] */
  const int32_t* aln_seq_off;     /* [n_alns+1] offsets into bases / quals */
  const char* bases;
  const char* quals;
  const int32_t* aln_cigar_off;   /* [n_alns+1] offsets into cigar_type / cigar_len */
  const char* cigar_type;
  const int32_t* cigar_len;
  int32_t n_sets;
  const int32_t* set_off;         /* [n_sets+1] offsets into the SNP arrays */
  const uint32_t* snp_pos;        /* SNP::pos() (0-based), ascending and distinct within a set */
  const char* snp_base1;          /* SNP::base_one(): allele on the sample's first haplotype */
  const char* snp_base2;
} hipstr_snp_phasing_t;
hipstr_status_t hipstr_snp_phasing_batch_host(hipstr_ctx_t* ctx, const hipstr_snp_phasing_t* batch, double* log_p1,
                                              double* log_p2, int32_t* counts);

/* --- section 8(f) row 4, second slice: BAM ingestion and read filtering (host) ---
 * Everything between the BAM files and SNPBamProcessor::process_reads for ONE region (BamProcessor::process_regions,
 * src/bam_processor.cpp:551-607), on opaque handles:
 *   hipstr_bam_reader_*   BamCramMultiReader(paths, "", ORDER_ALNS_BY_FILE) + SetRegion + GetNextAlignment
 *                         (src/bam_io.{h,cpp}) for BAM files with a .bai index (BGZF, BAM and the index are decoded by
 *                         hipstr_b200/host/bam_reader.cpp; CRAM is out of scope)
 *   hipstr_filter_reads   BamProcessor::read_and_filter_reads (:173-474): quality trimming, adapter trimming, the N /
 *                         base-quality / end-match / indel-distance filters, mate pairing with the unique-mapping rules
 *                         (XA / SA / AS / XS tags), grouping by sample; then remove_pcr_duplicates
 *                         (src/pcr_duplicates.cpp) when options.remove_pcr_dups
 *   hipstr_filtered_reads_view   the kept reads as the flat arrays K7 takes (entries in the order of
 *                         SNPBamProcessor::process_reads: per sample the paired STR reads, each followed by its mate,
 *                         then the unpaired ones); entry_snp_set holds the sample index.
 * The *_text functions print one line per alignment (for inspection and for the parity tests); they return the text
 * length, or -(needed capacity) when cap is too small.  Errors the reference dies on return HIPSTR_ERR_BAD_ARG with
 * the message in hipstr_ingest_last_error() (thread-local). */
typedef struct hipstr_bam_reader hipstr_bam_reader_t;
typedef struct hipstr_bam_records hipstr_bam_records_t;
typedef struct hipstr_filtered_reads hipstr_filtered_reads_t;
typedef struct {                       /* BamProcessor's public knobs (src/bam_processor.h:78-101, 164-177) */
  int32_t max_mate_dist;               /* MAX_MATE_DIST            1000 */
  int32_t min_bp_before_indel;         /* MIN_BP_BEFORE_INDEL      7 */
  int32_t min_flank;                   /* MIN_FLANK                5 */
  int32_t min_read_end_match;          /* MIN_READ_END_MATCH       10 */
  int32_t maximal_end_match_window;    /* MAXIMAL_END_MATCH_WINDOW 15 */
  int32_t require_paired_reads;        /* REQUIRE_PAIRED_READS     1 */
  double  min_sum_qual_log_prob;       /* MIN_SUM_QUAL_LOG_PROB    -10 */
  int32_t max_total_reads;             /* MAX_TOTAL_READS          1000000 */
  int32_t base_qual_trim;              /* BASE_QUAL_TRIM           '5' (character code) */
  int32_t remove_pcr_dups;             /* REMOVE_PCR_DUPS          1 */
  int32_t trim_adapters;               /* AdapterTrimmer on (TruSeq + Nextera), 1 */
} hipstr_filter_options_t;
typedef struct {
  int32_t n_samples;
  const char* const* sample_names;     /* rg_names, in order of appearance */
  const int32_t* sample_entry_off;     /* [n_samples+1] entries of sample s */
  const char* entry_passes;            /* [n_entries] '1' if the STR read may be used to generate haplotypes (PF tag, first region) */
  const int32_t* aln_flag;             /* [n_alns] BAM FLAG */
  const int32_t* entry_name_off;       /* [n_entries+1] into entry_names */
  const char* entry_names;             /* BamAlignment::Name() of the STR reads (adjacent equal names = mates that both span the STR) */
  hipstr_snp_phasing_t reads;          /* SNP arrays left NULL / 0: the caller attaches its SNP sets */
} hipstr_filtered_view_t;
const char* hipstr_ingest_last_error(void);
hipstr_status_t hipstr_bam_reader_open(int32_t n_files, const char* const* paths, hipstr_bam_reader_t** out);
void hipstr_bam_reader_close(hipstr_bam_reader_t* reader);
/* "@RG" lines of every file (BamHeader::parse_read_groups): path, ID, SM, LB per line, "-" when a tag is absent */
int64_t hipstr_bam_reader_read_groups(const hipstr_bam_reader_t* reader, int64_t cap, char* out_text);
/* the alignments overlapping [start, end) of `chrom`, file after file */
hipstr_status_t hipstr_bam_reader_fetch(hipstr_bam_reader_t* reader, const char* chrom, int32_t start, int32_t end,
                                        hipstr_bam_records_t** out);
int32_t hipstr_bam_records_count(const hipstr_bam_records_t* records);
int64_t hipstr_bam_records_text(const hipstr_bam_records_t* records, int64_t cap, char* out_text);
void hipstr_bam_records_free(hipstr_bam_records_t* records);
void hipstr_filter_default_options(hipstr_filter_options_t* options);
/* records: hipstr_bam_reader_fetch(chrom, region start - max_mate_dist (>= 0), region stop + max_mate_dist);
 * rg_keys[i] = file path + read group ID -> rg_samples[i] / rg_libraries[i] (what main builds from the headers) */
hipstr_status_t hipstr_filter_reads(const hipstr_bam_records_t* records, const char* chrom_seq, int32_t n_regions,
                                    const int32_t* region_start, const int32_t* region_stop,
                                    const hipstr_filter_options_t* options, int32_t n_rg, const char* const* rg_keys,
                                    const char* const* rg_samples, const char* const* rg_libraries,
                                    hipstr_filtered_reads_t** out);
/* counts9: reads overlapping the region, hard clipped, with an N, with low base qualities, without a unique mapping,
 * without a mate, sets of PCR duplicates removed, TOO_MANY_READS, reads that passed */
void hipstr_filtered_reads_counts(const hipstr_filtered_reads_t* reads, int32_t* counts9);
int64_t hipstr_filtered_reads_text(const hipstr_filtered_reads_t* reads, int64_t cap, char* out_text);
hipstr_status_t hipstr_filtered_reads_view(hipstr_filtered_reads_t* reads, hipstr_filtered_view_t* view);
void hipstr_filtered_reads_free(hipstr_filtered_reads_t* reads);
/* One alignment through one step, for checking without files: what = 0 BamAlignment::TrimLowQualityEnds((char)arg),
 * 1 AdapterTrimmer::trim_adapters (flag decides the mate / strand), 2 TrimNumBases(arg, arg2).  out_pos = {Position(),
 * GetEndPosition(), Length()}.  Returns 0, or -1 where the reference would have died. */
int32_t hipstr_trim_one(int32_t what, int32_t arg, int32_t arg2, int32_t flag, int32_t pos, int32_t end_pos, const char* bases,
                        const char* quals, int32_t n_cigar, const char* cigar_type, const int32_t* cigar_len, int32_t* out_pos,
                        char* out_seq, char* out_qual, int32_t* n_out_cigar, char* out_ctype, int32_t* out_clen);
/* out = {HasLargestEndMatches(aln, chrom_seq, 0, window, window), GetNumEndMatches first, second, GetEndDistToIndel first,
 * second} (src/alignment_filters.cpp), sum_qual = BaseQuality::sum_log_prob_correct(Qualities()) */
int32_t hipstr_alignment_filters(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar,
                                 const char* cigar_type, const int32_t* cigar_len, const char* chrom_seq, int32_t window,
                                 int32_t* out, double* sum_qual);

/* --- section 8(f) row 4, third slice: the phased SNP VCF behind K7 (host) ---------
 * Replaces VCF::VCFReader / VCF::Variant (src/vcf_reader.{h,cpp}) as create_snp_trees uses them (src/snp_tree.cpp:26-108):
 * the file (bgzipped or plain text) is read once, and a region query returns, for every sample of the VCF, the SNP set
 * create_snp_trees would put into that sample's SNPTree -- biallelic SNP records with start <= POS <= end (the tabix
 * region "chrom:start-end"; process_reads passes region start - MAX_MATE_DIST (or 1) and region stop + MAX_MATE_DIST,
 * src/snp_bam_processor.cpp:62), not within skip_padding (SKIP_PADDING = 15) of a region to skip, and of those the
 * sample's phased heterozygous calls as (POS - 1, allele on haplotype one, allele on haplotype two).
 * *found = 0 when the chromosome is not in the VCF (the reference then proceeds without SNP information).
 * The returned arrays (n_samples + 1 offsets, then the SNPs sample after sample) are K7's set_off / snp_pos / snp_base1 /
 * snp_base2 and stay valid until the next call on the handle.  Pedigree filtering is not built. */
typedef struct hipstr_snp_vcf hipstr_snp_vcf_t;
const char* hipstr_snp_vcf_last_error(void);
hipstr_status_t hipstr_snp_vcf_open(const char* path, hipstr_snp_vcf_t** out);
void hipstr_snp_vcf_close(hipstr_snp_vcf_t* vcf);
int32_t hipstr_snp_vcf_num_samples(const hipstr_snp_vcf_t* vcf);
const char* hipstr_snp_vcf_samples(const hipstr_snp_vcf_t* vcf);          /* sample names, one per line */
int32_t hipstr_snp_vcf_has_chromosome(const hipstr_snp_vcf_t* vcf, const char* chrom);
hipstr_status_t hipstr_snp_vcf_region_sets(hipstr_snp_vcf_t* vcf, const char* chrom, int32_t start, int32_t end, int32_t n_skip,
                                           const int32_t* skip_start, const int32_t* skip_stop, int32_t skip_padding,
                                           int32_t* found, const int32_t** set_off, const uint32_t** snp_pos,
                                           const char** snp_base1, const char** snp_base2);

/* The reference panel (--ref-vcf): read_vcf_alleles (src/vcf_input.cpp:21-50) -- among the records overlapping the region
 * padded by 50 bp, the first whose INFO START / END equal region start + 1 / region stop gives *pos (0-based POS) and the
 * alleles (REF first, one per line in *alleles_text, valid until the next call); returns 1, or 0 when there is none. */
typedef struct hipstr_str_vcf hipstr_str_vcf_t;
hipstr_status_t hipstr_str_vcf_open(const char* path, hipstr_str_vcf_t** out);
void hipstr_str_vcf_close(hipstr_str_vcf_t* vcf);
int32_t hipstr_str_vcf_alleles(hipstr_str_vcf_t* vcf, const char* chrom, int32_t region_start, int32_t region_stop, int32_t* pos,
                               int32_t* n_alleles, const char** alleles_text);

/* --- the per-region driver: BAM files -> VCF records for a WINDOW of regions -----
 * Replaces, for BAM input, BamProcessor::process_regions (src/bam_processor.cpp:521-617) -> SNPBamProcessor::process_reads
 * (src/snp_bam_processor.cpp:36-118) -> GenotyperBamProcessor::analyze_reads_and_phasing / learn_stutter_model
 * (src/genotyper_bam_processor.cpp:104-289) with the same decisions, re-arranged around window-sized device calls
 * (hipstr_b200/host/region_driver.cpp): regions are read and filtered on all host threads, then ONE K7 launch (when a
 * SNP VCF is given), ONE K4 call (unless use_def_stutter_model), the K6 left alignment of all reads and the lockstep
 * genotyper (K1 K2 K3 K5, then K3b K5 for the records) run over all regions of the window together.
 * Regions come in the order of the region file after orderRegions (src/region.cpp:53-55); chromosome sequences by name
 * (upper or lower case, as FastaReader returns them).  The status of a region is 0 genotyped (record available),
 * 1 longer than max_str_length, 2 within 50 bp of a contig end, 3 fewer than min_total_reads (informative) reads,
 * 4 too many reads, 5 stutter model training failed, 6 genotyping failed, 7 chromosome sequence not supplied. */
typedef struct {
  hipstr_filter_options_t filter;      /* BamProcessor's read-filter knobs */
  int32_t max_str_length;              /* MAX_STR_LENGTH       100 */
  int32_t min_total_reads;             /* MIN_TOTAL_READS      100 */
  int32_t max_total_haplotypes;        /* MAX_TOTAL_HAPLOTYPES 1000 */
  int32_t max_flank_haplotypes;        /* MAX_FLANK_HAPLOTYPES 4 */
  double  min_flank_freq;              /* MIN_FLANK_FREQ       0.01 */
  int32_t max_em_iter;                 /* MAX_EM_ITER          100 */
  double  abs_ll_converge;             /* ABS_LL_CONVERGE      0.01 */
  double  frac_ll_converge;            /* FRAC_LL_CONVERGE     0.001 */
  int32_t use_def_stutter_model;       /* --def-stutter-model: use def_stutter_model instead of EM training */
  double  def_stutter_model[6];        /* 0.95 0.05 0.05 0.95 0.01 0.01 (hipstr_main.cpp:343) */
  int32_t recalc_stutter_model;        /* recalc_stutter_model_ */
  int32_t skip_padding;                /* SNPBamProcessor::SKIP_PADDING 15 */
  int32_t n_haploid_chroms;            /* --haploid-chrs */
  const char* const* haploid_chroms;
  int32_t host_threads;                /* 0 = HIPSTR_HOST_THREADS / all cores */
  int32_t bams_from_10x;               /* --10x-bams: phasing from the reads' HP tags (SNPBamProcessor::process_10x_reads,
                                          src/snp_bam_processor.cpp:140-200) instead of a SNP VCF */
  struct hipstr_str_vcf* ref_vcf;      /* --ref-vcf: genotype the alleles of this reference panel (hipstr_str_vcf_open) instead of
                                          alleles found in the reads; NULL = none */
} hipstr_pipeline_options_t;
typedef struct hipstr_region_results hipstr_region_results_t;
void hipstr_pipeline_default_options(hipstr_pipeline_options_t* options);
const char* hipstr_process_regions_last_error(void);
hipstr_status_t hipstr_process_regions(hipstr_ctx_t* ctx, int32_t n_files, const char* const* bam_paths, hipstr_snp_vcf_t* snp_vcf,
                                       int32_t n_chroms, const char* const* chrom_names, const char* const* chrom_seqs,
                                       int32_t n_regions, const char* const* region_chrom, const int32_t* region_start,
                                       const int32_t* region_stop, const int32_t* region_period, const char* const* region_name,
                                       const hipstr_pipeline_options_t* options, const hipstr_vcf_options_t* vcf_options,
                                       hipstr_region_results_t** out);
int32_t hipstr_region_results_count(const hipstr_region_results_t* results);
/* status of a region (see above); *pos = POS of its record, *n_reads = reads that passed the filters */
int32_t hipstr_region_results_status(const hipstr_region_results_t* results, int32_t region, int32_t* pos, int32_t* n_reads);
const char* hipstr_region_results_record(const hipstr_region_results_t* results, int32_t region);   /* "" unless status 0 */
const char* hipstr_region_results_samples(const hipstr_region_results_t* results);   /* the VCF's sample columns, one per line */
/* seconds8: ingestion (all threads, wall), SNP sets + K7, stutter models, left alignment, genotyping, records, and of the
 * second: packing the window's reads and SNP sets, the K7 call itself;
 * counters4: alignments read, reads kept, reads with phase information, reads that failed to left-align */
void hipstr_region_results_timing(const hipstr_region_results_t* results, double* seconds8, int64_t* counters4);
/* hipstr_genotyper_timing / hipstr_genotyper_stats of the window's genotyper: seconds9 as there, stats3 = alignments, traces, rounds */
void hipstr_region_results_genotyper_timing(const hipstr_region_results_t* results, double* seconds9, int64_t* stats3);
void hipstr_region_results_free(hipstr_region_results_t* results);

/* Wall-clock seconds this context has spent inside hipstr_trace_batch_host, by part:
 * {host lowering of the batch, ordering + uploads, kernel K5, downloads of the results} */
void hipstr_trace_seconds(const hipstr_ctx_t* ctx, double* seconds4);

/* Accounting of the last public call on this context: bytes copied host->device and
 * device->host, and kernels launched. */
void hipstr_last_traffic(const hipstr_ctx_t* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes,
                         int32_t* kernel_launches);

#ifdef __cplusplus
}
#endif
#endif /* HIPSTR_B200_H_ */
