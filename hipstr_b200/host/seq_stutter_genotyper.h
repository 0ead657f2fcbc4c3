/*
 * seq_stutter_genotyper.h -- host side of seam B1: the per-locus control loop of the
 * reference's SeqStutterGenotyper (src/seq_stutter_genotyper.h:143-196, .cpp:219-415, 486-671, 805-879),
 * re-designed for a GPU: instead of one object per locus that aligns, traces and prunes serially,
 * a GenotyperBatch holds MANY loci and advances all of them in lockstep rounds.  Every round is at
 * most three batched device calls through the C-ABI --
 *     hipstr_trace_batch_host     (K5: the traces the loci are waiting for)
 *     hipstr_genotype_batch_host  (K1+K2+K3: loci that gained haplotypes, masked to the new columns)
 *     hipstr_posteriors_host      (K3: loci that only lost alleles)
 * -- and the host work in between is the reference's own decision logic per locus
 * (stutter-allele discovery, removal of uncalled / unspanned alleles, haplotype remapping by
 * sequence identity).  There is no CPU alignment path: without a context nothing here can run.
 *
 * Names follow the reference (SeqStutterGenotyper, HapBlock, genotype(), add_and_remove_alleles,
 * get_unused_alleles, get_stutter_candidate_alleles, retrace_alignments, trace_cache_).
 */
#ifndef HIPSTR_B200_SEQ_STUTTER_GENOTYPER_H_
#define HIPSTR_B200_SEQ_STUTTER_GENOTYPER_H_

#include <stdint.h>

#include <functional>
#include <map>
#include <unordered_map>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "../../include/hipstr_b200.h"

namespace hipstr {

/* HapBlock / RepeatBlock (SeqAlignment/HapBlock.h:18-148, RepeatBlock.h:15-70) as plain data. */
struct HapBlock {
  int32_t start = 0, end = 0;        /* reference coordinates, end exclusive */
  int32_t period = 0;                /* 0 = flank block, > 0 = repeat block with this motif length */
  double stutter[6] = {0, 0, 0, 0, 0, 0};
  std::vector<std::string> seqs;     /* [0] = reference allele, then alternates in insertion order */
  int num_options() const { return (int)seqs.size(); }
  bool contains(const std::string& s) const;
  HapBlock remove_alleles(const std::vector<int>& allele_indices) const;   /* HapBlock.h:138-147 */
};

/* What the control loop and the VCF writer read off an AlignmentTrace (AlignmentTraceback.h:9-117). */
struct AlignmentTrace {
  std::string hap_aln;                              /* hap_aln(): read vs haplotype operations (kept when keep_traced_alignments) */
  int32_t start = 0, stop = 0;                      /* traced_aln().get_start() / get_stop() */
  std::string cigar;                                /* traced_aln().getCigarString()   (filled when keep_traced_alignments) */
  std::string alignment;                            /* traced_aln().get_alignment()    (filled when keep_traced_alignments) */
  int32_t flank_ins_size = 0, flank_del_size = 0;
  /* Per block, in fixed slots (a locus has at most HIPSTR_MAX_BLOCKS_PER_LOCUS blocks): a trace is built for every
   * (pooled read, haplotype) the loop looks at, ~1000 per locus, so it owns no per-block heap storage. */
  int32_t num_blocks = 0;
  int32_t stutter_size[HIPSTR_MAX_BLOCKS_PER_LOCUS] = {};   /* HIPSTR_NO_STR_DATA where no STR data */
  /* the read bases aligned to the block: str_seq() of a repeat block, flank_seq() of a flank block -- views into the
   * locus' pooled read bases (SeqStutterGenotyper::pool_bases_), valid while the locus lives */
  std::string_view block_seq[HIPSTR_MAX_BLOCKS_PER_LOCUS];
  std::string_view str_seq(int block) const { return block_seq[block]; }
  std::string_view flank_seq(int block) const { return block_seq[block]; }
  std::vector<std::pair<int32_t, int32_t> > flank_indel_data;
  std::vector<std::pair<int32_t, char> > flank_snp_data;
  bool has_stutter() const {
    for (int b = 0; b < num_blocks; b++)
      if (stutter_size[b] != HIPSTR_NO_STR_DATA && stutter_size[b] != 0) return true;
    return false;
  }
  int total_stutter_size() const {
    int total = 0;
    for (int b = 0; b < num_blocks; b++)
      if (stutter_size[b] != HIPSTR_NO_STR_DATA) total += stutter_size[b];
    return total;
  }
};

/* Haplotype::aln_haps_to_ref for one haplotype (SeqAlignment/Haplotype.cpp:8-86): global affine-gap
 * alignment of the alternate haplotype against the reference haplotype (NeedlemanWunsch::Align with
 * the reference end penalty, NeedlemanWunsch.cpp:84-131,193-241,253-337,339-423), indels in the
 * upstream flank shifted right to the repeat block, encoded as one of 'M','I','D' per column. */
std::string hap_aln_to_ref(const std::string& ref_hap, const std::string& alt_hap, int32_t first_block_start,
                           int32_t repeat_block_start);

/* trace_cache_ (seq_stutter_genotyper.h:61 of the reference: std::map<std::pair<int,int>, AlignmentTrace*>): (pooled read,
 * haplotype) -> trace.  Every decision of the loop looks its reads up here, ~10 lookups per read and round, so the map is
 * a hash over one 64-bit key with the traces in a flat vector instead of a red-black tree of pairs. */
class TraceCache {
 public:
  typedef std::pair<int, int> Key;
  size_t count(const Key& k) const { return index_.count(pack(k)); }
  const AlignmentTrace& at(const Key& k) const { return items_[index_.at(pack(k))].second; }
  int32_t slot_of(const Key& k) const { return index_.at(pack(k)); }            /* position in insertion order */
  const AlignmentTrace& in_slot(int32_t i) const { return items_[i].second; }
  AlignmentTrace& operator[](const Key& k) {
    auto it = index_.find(pack(k));
    if (it != index_.end()) return items_[it->second].second;
    index_.emplace(pack(k), (int32_t)items_.size());
    items_.emplace_back(k, AlignmentTrace());
    return items_.back().second;
  }
  void clear() { index_.clear(); items_.clear(); }
  void swap(TraceCache& o) { index_.swap(o.index_); items_.swap(o.items_); }
  size_t size() const { return items_.size(); }
  void reserve(size_t n) { index_.reserve(n); items_.reserve(n); }
  std::vector<std::pair<Key, AlignmentTrace> >::iterator begin() { return items_.begin(); }
  std::vector<std::pair<Key, AlignmentTrace> >::iterator end() { return items_.end(); }
 private:
  static uint64_t pack(const Key& k) { return ((uint64_t)(uint32_t)k.first << 32) | (uint32_t)k.second; }
  std::unordered_map<uint64_t, int32_t> index_;
  std::vector<std::pair<Key, AlignmentTrace> > items_;
};

class GenotyperBatch;
double now_s();   /* steady clock, seconds */
/* Runs fn(i) for i in [0, n) on the host cores (std::thread, dynamic scheduling).  The per-locus host logic of the
 * loop is independent across loci; thread count = HIPSTR_HOST_THREADS or the hardware concurrency (at most 32). */
int host_threads();
void parallel_for(size_t n, const std::function<void(size_t)>& fn);

/* One locus.  Member names follow seq_stutter_genotyper.h:28-68 / genotyper.h:20-46. */
class SeqStutterGenotyper {
 public:
  enum Phase { ALIGN_ALL, STUTTER_ALLELES, PRUNE_UNCALLED, PRUNE_UNSPANNED, ASSEMBLE_FLANKS, ASSEMBLE_PRUNE, DONE, FAILED };
  enum Request { NONE, NEED_TRACES, NEED_ALIGNMENT, NEED_POSTERIORS };

  bool haploid_ = false;
  int num_samples_ = 0, num_reads_ = 0, num_alleles_ = 0;   /* num_alleles_ = number of haplotypes */
  std::vector<HapBlock> hap_blocks_;
  std::vector<std::string> hap_aln_info_;                    /* per haplotype, Haplotype::get_aln_info */
  std::vector<std::vector<int32_t> > hap_aln_index_;         /* per haplotype, hipstr_hap_aln_index of the string above */
  /* reads (sample-major) */
  std::vector<int32_t> sample_label_, pool_index_, read_weights_, seed_positions_;
  std::vector<uint8_t> second_mate_;
  std::vector<double> log_p1_, log_p2_;
  std::vector<int32_t> read_start_, read_cigar_off_, read_cigar_len_;
  std::vector<char> read_cigar_type_;
  std::vector<uint8_t> rev_strand_;            /* Alignment::is_from_reverse_strand() */
  std::vector<int32_t> read_seq_off_;          /* [R+1] into read_quals_ */
  std::string read_quals_;                     /* the reads' own base qualities (pools carry the medians) */
  /* pooled reads (ReadPooler) */
  int num_pools_ = 0;
  std::vector<int32_t> pool_seq_off_, pool_seed_;
  std::string pool_bases_, pool_quals_;
  /* results */
  std::vector<double> log_aln_probs_;          /* [R][H] */
  std::vector<double> log_sample_posteriors_;  /* [S][H][H] */
  std::vector<double> sample_total_LLs_;       /* [S] */
  std::vector<int32_t> optimal_haps_;          /* [S][2] get_optimal_haplotypes */
  std::vector<std::string> call_sample_;       /* non-empty = sample not genotyped, with the reason */
  TraceCache trace_cache_;                     /* (pool, haplotype) -> trace */
  std::string log_;
  std::string vcf_record_;                     /* text of the last write_vcf_record */
  int32_t vcf_pos_ = 0;                        /* its POS */
  double phase_seconds_[8] = {0, 0, 0, 0, 0, 0, 0, 0};   /* host seconds spent deciding, by Phase */
  int rounds_ = 0;                             /* alignment rounds run (1 = no allele discovery) */

  /* write_vcf_record (.cpp:995-1510) in two steps around one batched trace call */
  struct ReadCall { int best_hap; int read_strand; double log_phase_one; bool unique; };
  std::vector<ReadCall> vcf_calls_;
  void vcf_prepare(const int32_t* best_hap);   /* per-read strand / haplotype assignment, lists missing traces */
  bool reassemble_flanks() const { return reassemble_flanks_; }

  Phase phase() const { return phase_; }
  bool succeeded() const { return phase_ == DONE; }
  std::string pool_read(int pool) const { return pool_bases_.substr(pool_seq_off_[pool], pool_seq_off_[pool + 1] - pool_seq_off_[pool]); }
  std::string_view pool_read_view(int pool) const {
    return std::string_view(pool_bases_).substr(pool_seq_off_[pool], pool_seq_off_[pool + 1] - pool_seq_off_[pool]);
  }
  void haps_to_alleles(int block_index, std::vector<int>& allele_indices) const;   /* .cpp:219-227 */
  std::string hap_seq(int hap) const;

 private:
  friend class GenotyperBatch;
  Phase phase_ = ALIGN_ALL;
  int max_total_haplotypes_ = 1000, max_flank_haplotypes_ = 4;
  double min_flank_freq_ = 0.01;
  bool reassemble_flanks_ = false;
  bool fixed_alleles_ = false;   /* ref_vcf_ != NULL: the alleles come from a reference panel and are never added to or pruned */
  /* pending device work */
  std::vector<uint8_t> realign_hap_, realign_pool_, copy_read_;   /* masks of the pending alignment */
  std::vector<std::pair<int, int> > missing_traces_;
  /* write_vcf_record traces a read with ITS OWN qualities when its (pool, haplotype) trace is not cached
   * (.cpp:1118 passes alns_[read_index], retrace_alignments :831 the pooled read): the read whose qualities
   * to use per missing trace, empty = pooled qualities */
  std::vector<int> missing_trace_read_;

  Request advance();                                   /* runs host logic until device work is needed */
  int best_hap_of_read(int read) const;                /* retrace_alignments, .cpp:825-827 */
  bool collect_missing_traces();                       /* true if every needed trace is cached */
  void get_stutter_candidate_alleles(int block_index, std::vector<std::string>& candidate_seqs);
  void get_unused_alleles(bool check_spanned, bool check_called, std::vector<std::vector<int> >& allele_indices,
                          int& num_aff_blocks, int& num_aff_alleles);
  bool add_and_remove_alleles(const std::vector<std::vector<int> >& alleles_to_remove,
                              const std::vector<std::vector<std::string> >& alleles_to_add,
                              const std::vector<uint8_t>* realign_pool = nullptr,
                              const std::vector<uint8_t>* copy_read = nullptr);   /* true if K1 is needed */
  /* assemble_flanks (.cpp:40-217) up to the realignment request: 0 = nothing to realign, 1 = alignment
   * requested, -1 = the reference returns false (locus skipped) */
  int assemble_flanks();
  void rebuild_hap_aln_info(const std::map<std::string, std::string>* known);
};

/* The batch engine: seam B1 for many loci at once. */
class GenotyperBatch {
 public:
  explicit GenotyperBatch(hipstr_ctx_t* ctx) : ctx_(ctx) {}
  /* Mirrors the SeqStutterGenotyper constructor + init() (.cpp:486-517) for pre-built haplotype blocks:
   * pools the reads, marks second mates, computes pool seeds. */
  hipstr_status_t add_loci(const hipstr_align_batch_t* blocks, const int32_t* block_start, const int32_t* block_end,
                           const hipstr_locus_reads_t* reads, std::string& err);
  /* The whole constructor: build_haplotype (.cpp:422-484, HaplotypeGenerator) from the reads of one STR region per
   * locus, then init().  A locus whose haplotype construction fails stays uninitialised (genotype() = false). */
  hipstr_status_t add_loci_from_reads(int32_t n_loci, const int32_t* region_start, const int32_t* region_stop,
                                      const int32_t* period, const char* const* chrom_seq, const double* stutter,
                                      const hipstr_locus_reads_t* reads, std::string& err, const int32_t* allele_pos = nullptr,
                                      const int32_t* allele_off = nullptr, const char* const* alleles = nullptr);
  /* genotype() of every locus (.cpp:603-671), lockstep rounds. */
  hipstr_status_t genotype(int max_total_haplotypes, int max_flank_haplotypes, double min_flank_freq, bool reassemble_flanks,
                           std::string& err);

  /* recompute_stutter_models (.h:195-196, .cpp:1541-1583) of every genotyped locus: the STR sizes observed in the
   * maximum-likelihood alignments (len(str_seq) + stutter size of every spanning read) train a new stutter model per
   * repeat block -- one batched K4 call for all loci -- then genotype() runs again with it.  A locus whose training does
   * not converge fails, like the reference's `return false`. */
  hipstr_status_t recompute_stutter_models(int max_total_haplotypes, int max_flank_haplotypes, double min_flank_freq, int max_em_iter,
                                           double abs_ll_converge, double frac_ll_converge, std::string& err);

  std::vector<SeqStutterGenotyper> loci;
  bool keep_traced_alignments = false;   /* also build AlignmentTrace::cigar / alignment (only the visualisation reads them) */
  int64_t n_alignments = 0, n_traces = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0, gpu_launches = 0;   /* summed over every device call of the loop (hipstr_last_traffic) */
  void account_device_call();
  int n_rounds = 0;
  /* wall-clock seconds by stage: construction, per-locus host decisions, trace device calls, trace stitching +
   * bookkeeping, alignment calls (packing + K1/K2/K3 + unpacking), posterior calls, VCF formatting */
  enum { T_CONSTRUCT, T_DECIDE, T_TRACE_DEVICE, T_TRACE_HOST, T_ALIGN, T_POSTERIORS, T_VCF, T_ALIGN_PACK, T_ALIGN_UNPACK, T_COUNT };
  double seconds[T_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   /* the last two split T_ALIGN's host share: packing, unpacking */
  hipstr_status_t run_traces(const std::vector<int>& which, std::string& err);
  /* write_vcf_record of every successfully genotyped locus (seq_stutter_genotyper.h:179-181, impl .cpp:984-1510,
   * get_alleles :691-769, reorder_alleles :673-689, compute_allele_bias :965-982): K3b marginalises the posteriors of
   * all loci in one call, K5 traces the reads whose strand-assigned haplotype is not cached yet, then each record is
   * formatted on the host.  Records are kept in vcf_record_ / vcf_pos_ of each locus. */
  hipstr_status_t write_vcf_records(const hipstr_vcf_loci_t* regions, const hipstr_vcf_options_t* options, std::string& err);

 private:
  hipstr_ctx_t* ctx_;
  hipstr_status_t init_reads(SeqStutterGenotyper& g, const hipstr_locus_reads_t* reads, int locus, std::string& err);
  hipstr_status_t run_alignments(const std::vector<int>& which, std::string& err);
  hipstr_status_t run_posteriors(const std::vector<int>& which, std::string& err);
};

}  // namespace hipstr

/* the opaque handle of the C-ABI */
struct hipstr_genotyper {
  hipstr::GenotyperBatch batch;
  std::string last_error;
  explicit hipstr_genotyper(hipstr_ctx_t* ctx) : batch(ctx) {}
};
#endif
