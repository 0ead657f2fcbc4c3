/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" adapter that drives the UNMODIFIED reference classes
 * (compiled from /root/reference/src by oracle/Makefile, objects and the
 * resulting libhipstr_ref.so live only under oracle/_ref/) with the same flat
 * inputs as include/hipstr_b200.h.  No reference source is copied: this file
 * only #includes the reference headers from where they lie and calls the
 * public entry points the survey lists as seams B2-B4:
 *   HapAligner::process_read / calc_seed_base   (SeqAlignment/HapAligner.h:81-93)
 *   Genotyper::calc_log_sample_posteriors        (genotyper.h:72, via a test subclass)
 *   EMStutterGenotyper::train                    (em_stutter_genotyper.h:110)
 */
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "SeqAlignment/AlignmentData.h"
#include "SeqAlignment/AlignmentModel.h"
// the harness needs to see whether a block got STR data at all; AlignmentTrace has no accessor for that
#define private public
#include "SeqAlignment/AlignmentTraceback.h"
#undef private
#include "SeqAlignment/HapAligner.h"
#include "SeqAlignment/HapBlock.h"
#include "SeqAlignment/Haplotype.h"
#include "SeqAlignment/RepeatBlock.h"
#include "base_quality.h"
#include "em_stutter_genotyper.h"
#include "genotyper.h"
#include "mathops.h"
#include "stutter_model.h"

#include "../include/hipstr_b200.h"

namespace {

struct Init {
  Init() { precompute_integer_logs(); init_alignment_model(); }
};
void ensure_init() { static Init once; }

struct RefLocus {
  std::vector<HapBlock*> blocks;
  std::vector<StutterModel*> models;
  Haplotype* hap = NULL;
  RefLocus(const hipstr_align_batch_t* bt, int l, const int32_t* starts, const int32_t* ends) {
    int b0 = bt->locus_block_off[l], b1 = bt->locus_block_off[l + 1];
    int32_t pos = 0;
    for (int b = b0; b < b1; b++) {
      int o0 = bt->block_opt_off[b], o1 = bt->block_opt_off[b + 1];
      std::string ref(bt->opt_seq + bt->opt_seq_off[o0], bt->opt_seq + bt->opt_seq_off[o0 + 1]);
      int32_t st = starts ? starts[b - b0] : pos, en = ends ? ends[b - b0] : pos + (int32_t)ref.size();
      HapBlock* blk;
      if (bt->block_period[b] > 0) {
        const double* p = bt->block_stutter + 6 * (size_t)b;
        StutterModel* m = new StutterModel(p[0], p[1], p[2], p[3], p[4], p[5], bt->block_period[b]);
        models.push_back(m);
        blk = new RepeatBlock(st, en, ref, bt->block_period[b], m);
      } else
        blk = new HapBlock(st, en, ref);
      for (int o = o0 + 1; o < o1; o++)
        blk->add_alternate(std::string(bt->opt_seq + bt->opt_seq_off[o], bt->opt_seq + bt->opt_seq_off[o + 1]));
      blocks.push_back(blk);
      pos = en;
    }
    hap = new Haplotype(blocks);
  }
  ~RefLocus() {
    delete hap;
    for (auto b : blocks) delete b;
    for (auto m : models) delete m;
  }
};

class ExposedGenotyper : public Genotyper {
 public:
  ExposedGenotyper(bool haploid, const std::vector<std::string>& names, const std::vector<std::vector<double> >& p1,
                   const std::vector<std::vector<double> >& p2, int num_alleles)
      : Genotyper(haploid, names, p1, p2) {
    num_alleles_ = num_alleles;
    log_sample_posteriors_ = new double[num_samples_ * num_alleles_ * num_alleles_];
    log_aln_probs_ = new double[num_reads_ * num_alleles_];
  }
  void load_posteriors(const double* post, const double* sample_ll) {
    std::memcpy(log_sample_posteriors_, post, sizeof(double) * num_samples_ * num_alleles_ * num_alleles_);
    std::memcpy(sample_total_LLs_, sample_ll, sizeof(double) * num_samples_);
  }
  double run(const double* ll, const int32_t* weights, double* post, double* sample_ll, int32_t* best) {
    std::memcpy(log_aln_probs_, ll, sizeof(double) * num_reads_ * num_alleles_);
    std::vector<int> w(weights, weights + num_reads_);
    double total = calc_log_sample_posteriors(w);
    std::memcpy(post, log_sample_posteriors_, sizeof(double) * num_samples_ * num_alleles_ * num_alleles_);
    std::memcpy(sample_ll, sample_total_LLs_, sizeof(double) * num_samples_);
    if (best) {
      std::vector<std::pair<int, int> > gts;
      get_optimal_haplotypes(gts);
      for (int s = 0; s < num_samples_; s++) { best[2 * s] = gts[s].first; best[2 * s + 1] = gts[s].second; }
    }
    return total;
  }
};

}  // namespace

extern "C" {

double ref_fast_lse2(double a, double b) { return fast_log_sum_exp(a, b); }
double ref_fast_lse_vec(const double* v, int32_t n) {
  std::vector<double> x(v, v + n);
  return fast_log_sum_exp(x);
}

// Enumerate haplotypes with the reference iterator; out_opts is [n_haps][n_blocks].
int32_t ref_enumerate_haplotypes(const hipstr_align_batch_t* bt, int32_t locus, int32_t* out_opts) {
  ensure_init();
  RefLocus rl(bt, locus, NULL, NULL);
  int nb = rl.hap->num_blocks();
  int64_t h = 0;
  do {
    for (int b = 0; b < nb; b++) out_opts[h * nb + b] = rl.hap->cur_index(b);
    h++;
  } while (rl.hap->next());
  rl.hap->reset();
  return (int32_t)h;
}

int32_t ref_calc_seeds(int32_t n_reads, const int32_t* read_start, const int32_t* read_len, const int32_t* cigar_off,
                       const char* cigar_type, const int32_t* cigar_len, int32_t first_block_start,
                       int32_t last_block_end, int32_t n_repeats, const int32_t* repeat_start,
                       const int32_t* repeat_end, int32_t* out_seed) {
  ensure_init();
  // Dummy haplotype with the requested block coordinates: flank, (repeat, flank)*.
  std::vector<HapBlock*> blocks;
  StutterModel model(0.9, 0.01, 0.01, 0.9, 0.01, 0.01, 2);
  if (n_repeats != 1) return HIPSTR_ERR_UNSUPPORTED;  // Haplotype::adjust_indels asserts exactly 3 blocks
  blocks.push_back(new HapBlock(first_block_start, repeat_start[0], "ACGTAC"));
  blocks.push_back(new RepeatBlock(repeat_start[0], repeat_end[0], "ATATATAT", 2, &model));
  blocks.push_back(new HapBlock(repeat_end[0], last_block_end, "GATTAC"));
  Haplotype hap(blocks);
  std::vector<bool> mask(hap.num_combs(), true);
  {
    HapAligner aligner(&hap, mask);
    for (int r = 0; r < n_reads; r++) {
      std::string seq(read_len[r], 'A'), qual(read_len[r], 'I');
      int32_t ref_span = 0;
      Alignment aln(read_start[r], 0, false, "r", qual, seq, "");
      for (int c = cigar_off[r]; c < cigar_off[r + 1]; c++) {
        aln.add_cigar_element(CigarElement(cigar_type[c], cigar_len[c]));
        if (cigar_type[c] != 'I') ref_span += cigar_len[c];
      }
      aln.set_stop(read_start[r] + ref_span - 1);
      out_seed[r] = aligner.calc_seed_base(aln);
    }
  }
  for (auto b : blocks) delete b;
  return HIPSTR_OK;
}

int32_t ref_align_loci(const hipstr_align_batch_t* bt, int32_t l0, int32_t l1, double* ll_out, int32_t* seed_hap_pos) {
  ensure_init();
  (void)seed_hap_pos;
  BaseQuality base_quality;
  for (int l = l0; l < l1; l++) {
    RefLocus rl(bt, l, NULL, NULL);
    const int64_t H = rl.hap->num_combs();
    std::vector<bool> mask(H, true);
    if (bt->realign_hap)
      for (int64_t h = 0; h < H; h++) mask[h] = bt->realign_hap[bt->locus_hap_off[l] + h] != 0;
    HapAligner aligner(rl.hap, mask);
    AlignmentTrace trace(rl.hap->num_blocks());
    int p0 = bt->locus_pool_off[l], p1 = bt->locus_pool_off[l + 1];
    for (int p = p0; p < p1; p++) {
      if (bt->realign_pool && !bt->realign_pool[p]) continue;
      double* row = ll_out + bt->locus_out_off[l] + (int64_t)(p - p0) * H;
      int seed = bt->pool_seed[p];
      if (seed < 0) {
        for (int64_t h = 0; h < H; h++) row[h] = 0;
        continue;
      }
      int s0 = bt->pool_seq_off[p], s1 = bt->pool_seq_off[p + 1];
      Alignment aln(0, 0, false, "READPOOL", std::string(bt->pool_quals + s0, bt->pool_quals + s1),
                    std::string(bt->pool_bases + s0, bt->pool_bases + s1), "");
      aligner.process_read(aln, seed, &base_quality, false, row, trace);
    }
  }
  return HIPSTR_OK;
}

int32_t ref_align_batch(const hipstr_align_batch_t* bt, double* ll_out, int32_t* seed_hap_pos) {
  return ref_align_loci(bt, 0, bt->n_loci, ll_out, seed_hap_pos);
}

int32_t ref_posteriors(int32_t n_loci, const int32_t* locus_read_off, const int32_t* locus_sample_off,
                       const int32_t* n_haps, const uint8_t* haploid, const double* read_ll, const double* log_p1,
                       const double* log_p2, const int32_t* sample_label, const int32_t* read_weight,
                       double* post_out, double* sample_ll_out, int32_t* best_out, double* total_ll_out) {
  ensure_init();
  size_t ll_off = 0, post_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l], r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    const int s0 = locus_sample_off[l], S = locus_sample_off[l + 1] - s0;
    std::vector<std::string> names;
    std::vector<std::vector<double> > p1(S), p2(S);
    for (int s = 0; s < S; s++) { std::stringstream ss; ss << "S" << s; names.push_back(ss.str()); }
    int prev = 0;
    for (int r = r0; r < r1; r++) {
      if (sample_label[r] < prev) return HIPSTR_ERR_BAD_ARG;  // reads must be sample-major
      prev = sample_label[r];
      p1[sample_label[r]].push_back(log_p1[r]);
      p2[sample_label[r]].push_back(log_p2[r]);
    }
    ExposedGenotyper g(haploid[l] != 0, names, p1, p2, H);
    double total = g.run(read_ll + ll_off, read_weight + r0, post_out + post_off, sample_ll_out + s0,
                         best_out ? best_out + 2 * s0 : NULL);
    if (total_ll_out) total_ll_out[l] = total;
    ll_off += (size_t)(r1 - r0) * H;
    post_off += (size_t)S * H * H;
  }
  return HIPSTR_OK;
}

// HapAligner::trace_optimal_aln for a list of (pool, haplotype) pairs.  Besides the index-range outputs of
// hipstr_trace_out_t, the strings the reference keeps are returned verbatim so the tests can check that
// read[span] reproduces them: str_or_flank_seq is [n_traces][8][seq_stride].
int32_t ref_trace_batch(const hipstr_align_batch_t* bt, const int32_t* block_start, int32_t n_traces, const int32_t* trace_pool,
                        const int32_t* trace_hap, const hipstr_trace_out_t* out, char* str_or_flank_seq, int32_t seq_stride) {
  ensure_init();
  BaseQuality base_quality;
  for (int tr = 0; tr < n_traces; tr++) {
    const int p = trace_pool[tr];
    int l = 0;
    while (!(p >= bt->locus_pool_off[l] && p < bt->locus_pool_off[l + 1])) l++;
    const int b0 = bt->locus_block_off[l], nb = bt->locus_block_off[l + 1] - b0;
    std::vector<int32_t> starts(nb), ends(nb);
    for (int b = 0; b < nb; b++) {
      const int o0 = bt->block_opt_off[b0 + b];
      starts[b] = block_start[b0 + b];
      ends[b] = starts[b] + (bt->opt_seq_off[o0 + 1] - bt->opt_seq_off[o0]);
    }
    RefLocus rl(bt, l, starts.data(), ends.data());
    std::vector<bool> mask(rl.hap->num_combs(), true);
    HapAligner aligner(rl.hap, mask);
    const int s0 = bt->pool_seq_off[p], s1 = bt->pool_seq_off[p + 1];
    Alignment aln(0, 0, false, "READPOOL", std::string(bt->pool_quals + s0, bt->pool_quals + s1),
                  std::string(bt->pool_bases + s0, bt->pool_bases + s1), "");
    AlignmentTrace* trace = aligner.trace_optimal_aln(aln, bt->pool_seed[p], trace_hap[tr], &base_quality);
    const std::string& ha = trace->hap_aln();
    if ((int)ha.size() + 1 > out->aln_stride) { delete trace; return HIPSTR_ERR_BAD_ARG; }
    std::memcpy(out->hap_aln + (size_t)tr * out->aln_stride, ha.c_str(), ha.size() + 1);
    out->seed_hap_pos[tr] = -1;   // not observable through the public interface
    for (int b = 0; b < HIPSTR_MAX_BLOCKS_PER_LOCUS; b++) {
      const size_t o = (size_t)tr * HIPSTR_MAX_BLOCKS_PER_LOCUS + b;
      out->stutter_size[o] = HIPSTR_NO_STR_DATA; out->span_start[o] = 0; out->span_len[o] = 0;
      char* dst = str_or_flank_seq + o * seq_stride;
      dst[0] = 0;
      if (b >= nb) continue;
      std::string sq = trace->flank_seq(b);
      if (trace->str_data_[b] != NULL) { out->stutter_size[o] = trace->stutter_size(b); sq = trace->str_seq(b); }
      if ((int)sq.size() + 1 > seq_stride) { delete trace; return HIPSTR_ERR_BAD_ARG; }
      std::memcpy(dst, sq.c_str(), sq.size() + 1);
      out->span_len[o] = (int32_t)sq.size();
    }
    out->flank_ins[tr] = trace->flank_ins_size(); out->flank_del[tr] = trace->flank_del_size();
    const auto& ind = trace->flank_indel_data();
    const auto& snp = trace->flank_snp_data();
    out->n_indels[tr] = (int)ind.size(); out->n_snps[tr] = (int)snp.size();
    for (int k = 0; k < HIPSTR_MAX_TRACE_INDELS; k++) {
      const size_t o = ((size_t)tr * HIPSTR_MAX_TRACE_INDELS + k) * 2;
      out->indels[o] = k < (int)ind.size() ? ind[k].first : 0;
      out->indels[o + 1] = k < (int)ind.size() ? ind[k].second : 0;
    }
    for (int k = 0; k < HIPSTR_MAX_TRACE_SNPS; k++) {
      const size_t o = ((size_t)tr * HIPSTR_MAX_TRACE_SNPS + k) * 2;
      out->snps[o] = k < (int)snp.size() ? snp[k].first : 0;
      out->snps[o + 1] = k < (int)snp.size() ? (int)snp[k].second : 0;
    }
    delete trace;
  }
  return HIPSTR_OK;
}

// The COMPLETE flank_indel_data() / flank_snp_data() vectors of one trace (the batch entry above truncates them to the
// fixed slots of hipstr_trace_out_t): checker of hipstr_trace_flank_lists.
int32_t ref_trace_lists(const hipstr_align_batch_t* bt, const int32_t* block_start, int32_t pool, int32_t hap, int32_t cap_indels,
                        int32_t* n_indels, int32_t* indels, int32_t cap_snps, int32_t* n_snps, int32_t* snps) {
  ensure_init();
  BaseQuality base_quality;
  int l = 0;
  while (!(pool >= bt->locus_pool_off[l] && pool < bt->locus_pool_off[l + 1])) l++;
  const int b0 = bt->locus_block_off[l], nb = bt->locus_block_off[l + 1] - b0;
  std::vector<int32_t> starts(nb), ends(nb);
  for (int b = 0; b < nb; b++) {
    const int o0 = bt->block_opt_off[b0 + b];
    starts[b] = block_start[b0 + b];
    ends[b] = starts[b] + (bt->opt_seq_off[o0 + 1] - bt->opt_seq_off[o0]);
  }
  RefLocus rl(bt, l, starts.data(), ends.data());
  std::vector<bool> mask(rl.hap->num_combs(), true);
  HapAligner aligner(rl.hap, mask);
  const int s0 = bt->pool_seq_off[pool], s1 = bt->pool_seq_off[pool + 1];
  Alignment aln(0, 0, false, "READPOOL", std::string(bt->pool_quals + s0, bt->pool_quals + s1),
                std::string(bt->pool_bases + s0, bt->pool_bases + s1), "");
  AlignmentTrace* trace = aligner.trace_optimal_aln(aln, bt->pool_seed[pool], hap, &base_quality);
  const auto& ind = trace->flank_indel_data();
  const auto& snp = trace->flank_snp_data();
  *n_indels = (int32_t)ind.size();
  *n_snps = (int32_t)snp.size();
  for (int k = 0; k < (int)ind.size() && k < cap_indels; k++) { indels[2 * k] = ind[k].first; indels[2 * k + 1] = ind[k].second; }
  for (int k = 0; k < (int)snp.size() && k < cap_snps; k++) { snps[2 * k] = snp[k].first; snps[2 * k + 1] = (int)snp[k].second; }
  delete trace;
  return HIPSTR_OK;
}

// The reference's own stitched alignment of a trace (AlignmentTrace::traced_aln) together with the two inputs
// stitch_alignment_trace gets: hap_aln_to_ref = Haplotype::get_aln_info() of the traced haplotype (computed by the
// reference's Needleman-Wunsch in the Haplotype constructor) and the read-vs-haplotype string.
int32_t ref_trace_stitched(const hipstr_align_batch_t* bt, const int32_t* block_start, int32_t pool, int32_t hap,
                           char* hap_aln_to_ref, char* read_aln_to_hap, int32_t cap, int32_t* start, int32_t* stop,
                           char* cigar, char* alignment) {
  ensure_init();
  BaseQuality base_quality;
  int l = 0;
  while (!(pool >= bt->locus_pool_off[l] && pool < bt->locus_pool_off[l + 1])) l++;
  const int b0 = bt->locus_block_off[l], nb = bt->locus_block_off[l + 1] - b0;
  std::vector<int32_t> starts(nb), ends(nb);
  for (int b = 0; b < nb; b++) {
    const int o0 = bt->block_opt_off[b0 + b];
    starts[b] = block_start[b0 + b];
    ends[b] = starts[b] + (bt->opt_seq_off[o0 + 1] - bt->opt_seq_off[o0]);
  }
  RefLocus rl(bt, l, starts.data(), ends.data());
  rl.hap->go_to(hap);
  const std::string info = rl.hap->get_aln_info();
  rl.hap->reset();
  std::vector<bool> mask(rl.hap->num_combs(), true);
  HapAligner aligner(rl.hap, mask);
  const int s0 = bt->pool_seq_off[pool], s1 = bt->pool_seq_off[pool + 1];
  Alignment aln(0, 0, false, "READPOOL", std::string(bt->pool_quals + s0, bt->pool_quals + s1),
                std::string(bt->pool_bases + s0, bt->pool_bases + s1), "");
  AlignmentTrace* trace = aligner.trace_optimal_aln(aln, bt->pool_seed[pool], hap, &base_quality);
  Alignment& t = trace->traced_aln();
  std::stringstream cg;
  for (auto c = t.get_cigar_list().begin(); c != t.get_cigar_list().end(); c++) cg << c->get_num() << c->get_type();
  const std::string cgs = cg.str();
  if ((int)info.size() + 1 > cap || (int)trace->hap_aln().size() + 1 > cap || (int)cgs.size() + 1 > cap ||
      (int)t.get_alignment().size() + 1 > cap) { delete trace; return HIPSTR_ERR_BAD_ARG; }
  std::strcpy(hap_aln_to_ref, info.c_str());
  std::strcpy(read_aln_to_hap, trace->hap_aln().c_str());
  std::strcpy(cigar, cgs.c_str());
  std::strcpy(alignment, t.get_alignment().c_str());
  *start = t.get_start();
  *stop = t.get_stop();
  delete trace;
  return HIPSTR_OK;
}

// Genotyper::extract_genotypes_and_likelihoods for every locus.
int32_t ref_extract_genotypes(int32_t n_loci, const int32_t* locus_sample_off, const int32_t* n_haps, const int32_t* n_variants,
                              const int32_t* hap_to_allele, const uint8_t* haploid, const double* post, const double* sample_ll,
                              int32_t* best_hap, int32_t* best_gt, double* log_phased, double* log_unphased, double* hap_log_phased,
                              double* hap_log_unphased, double* gl, double* phased_gl, double* gl_diff, int32_t* pl) {
  ensure_init();
  size_t post_off = 0, h2a_off = 0, gl_off = 0, pgl_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l], V = n_variants[l], s0 = locus_sample_off[l], S = locus_sample_off[l + 1] - s0;
    const bool hap1 = haploid[l] != 0;
    const int G = hap1 ? V : V * (V + 1) / 2, PG = hap1 ? V : V * V;
    std::vector<std::string> names;
    std::vector<std::vector<double> > p1(S), p2(S);
    for (int s = 0; s < S; s++) { std::stringstream ss; ss << "S" << s; names.push_back(ss.str()); }
    ExposedGenotyper g(hap1, names, p1, p2, H);
    g.load_posteriors(post + post_off, sample_ll + s0);
    std::vector<int> h2a(hap_to_allele + h2a_off, hap_to_allele + h2a_off + H);
    std::vector<std::pair<int, int> > bh, bg;
    std::vector<double> lp, lu, hlp, hlu, gd;
    std::vector<std::vector<double> > gls, pgls;
    std::vector<std::vector<int> > pls;
    g.extract_genotypes_and_likelihoods(V, h2a, bh, bg, lp, lu, hlp, hlu, true, gls, gd, true, pls, true, pgls);
    for (int s = 0; s < S; s++) {
      best_hap[2 * (s0 + s)] = bh[s].first; best_hap[2 * (s0 + s) + 1] = bh[s].second;
      best_gt[2 * (s0 + s)] = bg[s].first; best_gt[2 * (s0 + s) + 1] = bg[s].second;
      log_phased[s0 + s] = lp[s]; log_unphased[s0 + s] = lu[s]; hap_log_phased[s0 + s] = hlp[s]; hap_log_unphased[s0 + s] = hlu[s];
      gl_diff[s0 + s] = gd[s];
      if ((int)gls[s].size() != G || (int)pgls[s].size() != PG) return HIPSTR_ERR_BAD_ARG;
      std::copy(gls[s].begin(), gls[s].end(), gl + gl_off + (size_t)s * G);
      std::copy(pgls[s].begin(), pgls[s].end(), phased_gl + pgl_off + (size_t)s * PG);
      std::copy(pls[s].begin(), pls[s].end(), pl + gl_off + (size_t)s * G);
    }
    post_off += (size_t)S * H * H; h2a_off += H; gl_off += (size_t)S * G; pgl_off += (size_t)S * PG;
  }
  return HIPSTR_OK;
}

// EMStutterGenotyper::train for every locus (seam B4).
int32_t ref_em_train(const hipstr_em_batch_t* bt, int32_t max_iter, double min_abs, double min_frac, double* params_out,
                     uint8_t* converged_out, int32_t* iters_out, double* ll_out) {
  ensure_init();
  (void)iters_out; (void)ll_out;   // not observable through the reference's public interface
  std::stringstream sink;
  for (int l = 0; l < bt->n_loci; l++) {
    const int r0 = bt->locus_read_off[l], r1 = bt->locus_read_off[l + 1];
    const int S = bt->locus_sample_off[l + 1] - bt->locus_sample_off[l];
    std::vector<std::vector<int> > bps(S);
    std::vector<std::vector<double> > p1(S), p2(S);
    std::vector<std::string> names;
    for (int s = 0; s < S; s++) { std::stringstream ss; ss << "S" << s; names.push_back(ss.str()); }
    for (int r = r0; r < r1; r++) {
      bps[bt->sample_label[r]].push_back(bt->num_bps[r]);
      p1[bt->sample_label[r]].push_back(bt->log_p1[r]);
      p2[bt->sample_label[r]].push_back(bt->log_p2[r]);
    }
    EMStutterGenotyper em(bt->haploid[l] != 0, bt->motif_len[l], bps, p1, p2, names, bt->ref_allele[l]);
    converged_out[l] = em.train(max_iter, min_abs, min_frac, false, sink);
    StutterModel* m = em.get_stutter_model();
    const char which[3] = {'P', 'U', 'D'};
    for (int k = 0; k < 6; k++) params_out[6 * (size_t)l + k] = m->get_parameter(k < 3, which[k % 3]);
  }
  return HIPSTR_OK;
}

}  // extern "C"
