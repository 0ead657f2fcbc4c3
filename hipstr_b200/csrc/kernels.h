/* kernels.h -- launchers of the sm_100a kernels (kernels.cu). */
#ifndef HIPSTR_B200_KERNELS_H_
#define HIPSTR_B200_KERNELS_H_
#include <cuda_runtime.h>
#include <stdint.h>
#include "layout.h"

namespace hipstr {

/* K1: read x haplotype HMM alignment.  `variant` indexes kColVariants (columns per lane). */
/* Persistent launch: at most max_ctas CTAs (the size of p.last_scratch in CTA slabs). */
#define HIPSTR_MAX_ALIGN_CTAS 8192
cudaError_t launch_align(int variant, const AlignParams& p, int max_ctas, cudaStream_t stream, int* grid_out);
size_t align_smem_bytes(int n_max, int l_max);
/* K1a: stutter tables of (pooled read, repeat allele) pairs (stutter.cu); runs before launch_align on the same stream. */
cudaError_t launch_stutter(const StutParams& p, cudaStream_t stream);
size_t stutter_smem_bytes(int n_max);

/* K2: pool -> read scatter + mate merge (seq_stutter_genotyper.cpp:530-564). */
struct ScatterParams {
  int32_t n_reads, n_haps;
  const double* pool_ll;
  const int32_t* pool_seed;
  const int32_t* pool_index;
  const uint8_t* second_mate;
  const uint8_t* copy_read;     /* may be NULL */
  const uint8_t* realign_hap;   /* may be NULL */
  double* read_ll;
  int32_t* read_seed;           /* may be NULL */
};
cudaError_t launch_scatter(const ScatterParams& p, cudaStream_t stream);

/* K2, batched over loci: same semantics, one launch for every locus of a batch. */
struct ScatterLocus {      /* one per locus */
  int64_t elem_off;        /* first (read, hap) element of the locus = read_ll offset */
  int64_t pool_ll_off;     /* locus_out_off: pool LL row 0 */
  int32_t read0;           /* global index of the locus's first read */
  int32_t n_reads;
  int32_t n_haps;
  int32_t pool0;           /* global index of the locus's first pool */
  int64_t hap0;            /* dense global index of haplotype 0 (mask lookup) */
};
struct ScatterBatchParams {
  int32_t n_loci;
  int64_t n_elems;
  const ScatterLocus* loci;
  const double* pool_ll;
  const int32_t* pool_seed;     /* global per pool */
  const int32_t* pool_index;    /* per read, local to locus */
  const uint8_t* second_mate;
  const uint8_t* copy_read;     /* may be NULL */
  const uint8_t* hap_mask;      /* may be NULL */
  double* read_ll;
  int32_t* read_seed;           /* may be NULL */
};
cudaError_t launch_scatter_batch(const ScatterBatchParams& p, cudaStream_t stream);

/* K3: genotype posteriors (genotyper.cpp:20-97). */
struct PostSample {       /* one per (locus, sample) */
  int32_t read0, read1;   /* global read range of the sample */
  int32_t n_haps;
  int32_t haploid;
  int64_t ll_off;         /* read_ll offset of the locus's first read (row = read - locus_read0) */
  int32_t locus_read0;
  int32_t locus;
  int64_t post_off;       /* post_out offset of this sample's H*H block */
};
struct PostParams {
  int32_t n_samples;      /* total over loci */
  int32_t n_loci;
  const PostSample* samples;
  const int32_t* locus_sample_off;
  const double* read_ll;
  const double* log_p1;
  const double* log_p2;
  const int32_t* read_weight;
  const double* int_logs;
  double log_one_half;
  double* post_out;
  double* sample_ll_out;
  int32_t* best_out;      /* may be NULL */
  double* total_ll_out;   /* may be NULL */
};
cudaError_t launch_posteriors(const PostParams& p, cudaStream_t stream);

/* K3b: genotypes and likelihoods from the posteriors (genotyper.cpp:99-251), one warp per (locus, sample). */
struct ExtractSample {     /* one per (locus, sample) */
  int32_t n_haps, n_variants, haploid;
  int32_t h2a_off;         /* into hap_to_allele */
  int64_t post_off;        /* this sample's H*H block */
  int64_t gl_off;          /* into gl / pl */
  int64_t pgl_off;         /* into phased_gl */
};
struct ExtractParams {
  int32_t n_samples;
  const ExtractSample* samples;
  const int32_t* hap_to_allele;
  const double* post;
  const double* sample_ll;
  const double* int_logs;
  int32_t* best_hap; int32_t* best_gt;
  double* log_phased; double* log_unphased; double* hap_log_phased; double* hap_log_unphased;
  double* gl; double* phased_gl; double* gl_diff; int32_t* pl;
};
cudaError_t launch_extract(const ExtractParams& p, cudaStream_t stream);

/* K4: EM stutter learner (em_stutter_genotyper.cpp:10-226), one CTA per locus for the whole training loop. */
struct EmLocus {            /* one per locus */
  int32_t read0, n_reads;   /* global read range */
  int32_t sample0, n_samples;
  int32_t n_alleles, period, haploid;
  int32_t allele_off;       /* into bps / gt_prior */
  int64_t post_off;         /* into post scratch: S * A * A doubles */
  int64_t row_off;          /* into row scratch: S * A doubles */
};
struct EmParams {
  int32_t n_loci, max_iter;
  double min_abs, min_frac;
  const EmLocus* loci;
  const int32_t* allele_of;       /* [R] allele index of each read */
  const int32_t* sample_label;    /* [R] sample of each read, local to its locus */
  const int32_t* sample_read_off; /* [S_total + 1] global read range of each sample */
  const double* log_p1;
  const double* log_p2;
  const int32_t* bps;             /* [sum A] bp size per allele (reference allele first) */
  double* gt_prior;               /* [sum A] in: init_log_gt_priors; updated in place */
  const double* int_logs;
  double log_one_half;
  double* post;                   /* scratch */
  double* rowlse;                 /* scratch */
  double* sample_ll;              /* scratch [S_total] */
  double* params;                 /* [n_loci][6] in: initial model; out: learned model */
  uint8_t* converged;             /* [n_loci] */
  int32_t* iters;                 /* [n_loci] */
  double* final_ll;               /* [n_loci] */
};
cudaError_t launch_em(const EmParams& p, int max_alleles, cudaStream_t stream);

/* K5: alignment traceback (trace.cu), one thread per (pooled read, haplotype) trace. */
struct TraceWalkParams {
  int32_t n_traces;
  const int32_t* trace_pool;     /* global pool index */
  const int32_t* trace_hap;      /* haplotype index local to the pool's locus */
  const DevPool* pools;
  const char* bases;
  const char* quals;
  const DevHapSide* hapsides;
  const uint8_t* hapbytes;
  const DevBlock* blocks;
  const double* qual_lut;
  const int32_t* block_start;    /* [n_blocks] genomic start of every block */
  const int32_t* block_ref_end;  /* [n_blocks] start + length of the reference allele */
  const int32_t* locus_block0;   /* [n_loci] first block of the locus */
  /* what the forward pass (k_align<.., TRACE>) left, see AlignParams */
  const unsigned char* dec; const int64_t* dec_off;
  const int32_t* art; const int64_t* art_off;
  const int32_t* seed_pos;
  int32_t aln_stride;
  char* out_aln;
  int32_t* out_seed_pos; int32_t* out_stutter; int32_t* out_span_start; int32_t* out_span_len;
  int32_t* out_flank_ins; int32_t* out_flank_del; int32_t* out_n_indels; int32_t* out_indels;
  int32_t* out_n_snps; int32_t* out_snps;
};
/* K5: forward pass (one warp per trace, kernels.cu) and walk back (one thread per trace, trace.cu) */
cudaError_t launch_trace_forward(int variant, const AlignParams& p, int max_ctas, cudaStream_t stream);
cudaError_t launch_trace_walk(const TraceWalkParams& p, cudaStream_t stream);

/* ---- K6: batched Needleman-Wunsch (nw.cu) -------------------------------------------------- */
struct NwParams {
  int32_t first_pair, n_pairs;  /* this launch handles pairs [first_pair, first_pair + n_pairs) */
  const int32_t* ref_off;    /* [total pairs+1] into ref_seqs */
  const char* ref_seqs;
  const int32_t* read_off;   /* [n_pairs+1] into read_seqs */
  const char* read_seqs;
  int32_t use_ref_end_penalty;
  int32_t max_ref, max_read; /* longest window / read of the batch (sizes the shared memory) */
  int32_t ops_stride;
  char* out_ops;             /* [n_pairs][ops_stride] 'M' / 'D' (reference base vs gap) / 'I' (read base vs gap), NUL-terminated */
  int32_t* out_len;          /* [n_pairs] number of operations, -1 if the walk back hit an impossible cell */
  float* out_score;          /* [n_pairs] */
  const int64_t* trace_off;  /* [total pairs+1] prefix sums of (window length rounded up to 16) x read length */
  unsigned char* trace;      /* trace bytes of the pairs of this launch, pair p at trace_off[p] - trace_off[first_pair] */
  int32_t* end_cell;         /* [total pairs][2] column and matrix where the alignment ends */
};
cudaError_t launch_nw(const NwParams& p, int max_ctas, cudaStream_t stream);
size_t nw_shared_bytes(int max_ref, int max_read);

/* ---- K7: SNP phasing log-likelihoods (snp_phase.cu) ---------------------------------------- */
struct SnpPhaseParams {
  int32_t n_entries;
  const int32_t* entry_aln_off;  /* [n_entries+1] alignments of an entry: the STR read, then its mate if any */
  const int32_t* entry_snp_set;  /* [n_entries] SNP set of the read's sample, -1 = none (both LLs stay 0) */
  const int32_t* aln_pos;        /* BamAlignment::Position() */
  const int32_t* aln_end;        /* BamAlignment::GetEndPosition(), exclusive */
  const int32_t* aln_seq_off;    /* [n_alns+1] into bases / quals */
  const char* bases;
  const char* quals;
  const int32_t* aln_cigar_off;  /* [n_alns+1] into cigar_type / cigar_len */
  const char* cigar_type;
  const int32_t* cigar_len;
  const int32_t* set_off;        /* [n_sets+1] into the SNP arrays; positions ascending within a set */
  const uint32_t* snp_pos;
  const char* snp_base1;
  const char* snp_base2;
  const double* qual_lut;        /* [256][2] */
  double* out_log_p1;
  double* out_log_p2;
  int32_t* out_counts;           /* [n_entries][4]: haplotype-1 matches, haplotype-2 matches, mismatches, status */
};
cudaError_t launch_snp_phase(const SnpPhaseParams& p, int n_sm, cudaStream_t stream);

}  // namespace hipstr

/* Launch helpers shared by the .cu files.  Several host threads (the pipelines of hipstr_multi_*) launch the same kernels with
 * different shared-memory sizes at the same time: setting cudaFuncAttributeMaxDynamicSharedMemorySize to the size of ONE launch
 * right before it races with the others (a launch then fails with "invalid argument" when another thread has just lowered the
 * limit).  The limit is therefore raised ONCE per (kernel, device) to the device's opt-in maximum; what a launch actually gets
 * is still its own `smem` argument. */
#ifdef __CUDACC__
#include <mutex>
#include <set>
#include <utility>
namespace hipstr {
inline cudaError_t allow_max_dynamic_smem(const void* kernel) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int> > done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(std::make_pair(kernel, dev))) return cudaSuccess;
  int optin = 0;
  if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  cudaFuncAttributes attr;
  if ((e = cudaFuncGetAttributes(&attr, kernel)) != cudaSuccess) return e;
  const int room = optin - (int)attr.sharedSizeBytes;   // the opt-in maximum covers static + dynamic shared memory
  if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, room)) != cudaSuccess) return e;
  done.insert(std::make_pair(kernel, dev));
  return cudaSuccess;
}
inline int sm_count_of_current_device() {
  static std::mutex mu;
  static int count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!count[dev]) cudaDeviceGetAttribute(&count[dev], cudaDevAttrMultiProcessorCount, dev);
  return count[dev];
}
}  // namespace hipstr
#endif

#endif
