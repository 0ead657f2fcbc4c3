/*
 * hipstr_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain serial C++) of the reference algorithm for the hot
 * path, taking the same flat inputs as the C-ABI in include/hipstr_b200.h.
 * It exists to check the CUDA path; nothing under hipstr_b200/ may call it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load liboracle.so.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md 4, 8c).  The restatement is pinned against the UNMODIFIED
 * reference sources compiled into oracle/_ref/libhipstr_ref.so (see
 * oracle/Makefile, oracle/ref_harness.cpp) by tests/test_oracle_vs_ref.py in
 * the build container, and against fixtures generated from that library and
 * committed under tests/golden/ (tests/golden/make_golden.py).
 */
#ifndef HIPSTR_ORACLE_H_
#define HIPSTR_ORACLE_H_

#include "../include/hipstr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The reference's two approximate log-sum-exp forms (mathops.cpp:86-106). */
double oracle_fast_lse2(double a, double b);
double oracle_fast_lse_vec(const double* v, int32_t n);

/* Per-block option index of haplotype `hap` (Haplotype.cpp:157-196). */
void oracle_hap_options(int32_t n_blocks, const int32_t* n_opts, int64_t hap, int32_t* out_opts);

int32_t oracle_calc_seeds(int32_t n_reads, const int32_t* read_start, const int32_t* read_len,
                          const int32_t* cigar_off, const char* cigar_type, const int32_t* cigar_len,
                          int32_t first_block_start, int32_t last_block_end,
                          int32_t n_repeats, const int32_t* repeat_start, const int32_t* repeat_end,
                          int32_t* out_seed);

int32_t oracle_align_batch(const hipstr_align_batch_t* batch, double* ll_out, int32_t* seed_hap_pos);

/* Same, restricted to loci [locus_begin, locus_end) -- lets bench.py fan the
 * CPU baseline out over host cores by forking over locus shards. */
int32_t oracle_align_loci(const hipstr_align_batch_t* batch, int32_t locus_begin, int32_t locus_end,
                          double* ll_out, int32_t* seed_hap_pos);

int32_t oracle_scatter_pool_lls(int32_t n_reads, int32_t n_haps, const double* pool_ll,
                                const int32_t* pool_seed, const int32_t* pool_index,
                                const uint8_t* second_mate, const uint8_t* copy_read,
                                const uint8_t* realign_hap, double* read_ll, int32_t* read_seed);

int32_t oracle_posteriors(int32_t n_loci, const int32_t* locus_read_off, const int32_t* locus_sample_off,
                          const int32_t* n_haps, const uint8_t* haploid, const double* read_ll,
                          const double* log_p1, const double* log_p2, const int32_t* sample_label,
                          const int32_t* read_weight, double* post_out, double* sample_ll_out,
                          int32_t* best_out, double* total_ll_out);

/* Genotyper::extract_genotypes_and_likelihoods (genotyper.cpp:129-251); same arguments as the product entry. */
int32_t oracle_extract_genotypes(int32_t n_loci, const int32_t* locus_sample_off, const int32_t* n_haps,
                                 const int32_t* n_variants, const int32_t* hap_to_allele, const uint8_t* haploid,
                                 const double* post, const double* sample_ll, int32_t* best_hap, int32_t* best_gt,
                                 double* log_phased, double* log_unphased, double* hap_log_phased,
                                 double* hap_log_unphased, double* gl, double* phased_gl, double* gl_diff, int32_t* pl);

/* HapAligner::trace_optimal_aln for a list of (pool, haplotype) pairs; same arguments as the product entry. */
int32_t oracle_trace_batch(const hipstr_align_batch_t* batch, const int32_t* block_start, int32_t n_traces,
                           const int32_t* trace_pool, const int32_t* trace_hap, const hipstr_trace_out_t* out);

/* EMStutterGenotyper::train for every locus of the batch (em_stutter_genotyper.cpp:170-226). */
/* NeedlemanWunsch::Align for one pair (checker of K6); ops capacity >= L1 + L2 + 1 */
int32_t oracle_nw_align(const char* ref, int32_t L1, const char* read, int32_t L2, int32_t use_ref_end_penalty, char* ops,
                        float* score);
int32_t oracle_em_train(const hipstr_em_batch_t* batch, int32_t max_iter, double min_LL_abs_change,
                        double min_LL_frac_change, double* params_out, uint8_t* converged_out,
                        int32_t* iters_out, double* ll_out);

/* calc_het_snp_factors for every entry of the batch (checker of K7); counts [n_entries][4] like the product;
 * returns 0, or < 0 where the reference would have died. */
int32_t oracle_snp_phasing(const hipstr_snp_phasing_t* batch, double* log_p1, double* log_p2, int32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
