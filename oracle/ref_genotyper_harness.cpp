/*
 * ref_genotyper_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" adapter around the UNMODIFIED reference SeqStutterGenotyper (seam B1,
 * src/seq_stutter_genotyper.h:143-196), compiled by oracle/Makefile from the sources where
 * they lie under /root/reference into oracle/_ref/libhipstr_ref.so.  It lets the tests run
 * the reference's whole per-locus control loop -- constructor (pooling + HaplotypeGenerator),
 * genotype() (align-all, posteriors, stutter-allele discovery rounds, allele pruning, flank
 * assembly) and write_vcf_record() -- on the same flat inputs the product takes, and read the
 * member arrays the product must reproduce.  `#define private public` is only used to READ
 * members (hap_blocks_, log_aln_probs_, ...); no reference source is modified or copied.
 *
 * kt_fisher_exact (htslib kfunc.c) and bdtr (cephes), which write_vcf_record reaches, come from the
 * vendored C files that oracle/Makefile compiles one by one (htslib 1.9 in full, see ref_bam_harness.cpp).
 */
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "vcf_writer.h"
#include "genotyper.h"
#include "seq_stutter_genotyper.h"
#undef private
#undef protected
#include "SeqAlignment/HapBlock.h"
#include "SeqAlignment/Haplotype.h"
#include "SeqAlignment/RepeatBlock.h"
#include "mathops.h"
#include "region.h"
#include "vcf_reader.h"
#include "stutter_model.h"
#include "extract_indels.h"
#include "debruijn_graph.h"
#include "SeqAlignment/NeedlemanWunsch.h"
#include "SeqAlignment/AlignmentOps.h"
#include "cephes/cephes.h"
#include "htslib/htslib/kfunc.h"
#include <cmath>
#include <algorithm>

namespace {

struct Init {
  Init() { precompute_integer_logs(); }
};
void ensure_init() { static Init once; }

struct RefSG {
  SeqStutterGenotyper* g = NULL;
  std::vector<StutterModel*> models;
  std::vector<std::string> names;
  std::string chrom_seq;
  std::ostringstream log;
  int n_reads = 0;
  VCF::VCFReader* ref_vcf = NULL;
  ~RefSG() {
    delete g;
    delete ref_vcf;
    for (auto m : models) delete m;
  }
};

}  // namespace

extern "C" {

/* Reads are sample-major; read r of sample s is named "r<name_id[r]>", so that adjacent reads with the
 * same id are mates (seq_stutter_genotyper.cpp:499).  The gapped alignment string is derived from bases +
 * CIGAR ('=','X','I','D'); read_stop is Alignment::get_stop() (HipSTR: last aligned reference position). */
void* ref_sg_create(int32_t n_samples, int32_t n_reads, const int32_t* sample_label, const int32_t* name_id,
                    const int32_t* read_start, const int32_t* read_stop, const int32_t* seq_off, const char* bases, const char* quals,
                    const int32_t* cigar_off, const char* cigar_type, const int32_t* cigar_len, const double* log_p1,
                    const double* log_p2, const char* chrom_seq, int32_t region_start, int32_t region_stop,
                    int32_t period, const double* stutter, int32_t haploid, int32_t reassemble_flanks,
                    const uint8_t* rev_strand, const uint8_t* use_for_haps, const char* ref_vcf_path /* NULL = no reference panel */) {
  ensure_init();
  RefSG* h = new RefSG();
  h->chrom_seq = chrom_seq;
  h->n_reads = n_reads;
  std::vector<Alignment> alns;
  std::vector<std::vector<double> > p1(n_samples), p2(n_samples);
  for (int s = 0; s < n_samples; s++) h->names.push_back("S" + std::to_string(s));
  for (int r = 0; r < n_reads; r++) {
    std::string seq(bases + seq_off[r], bases + seq_off[r + 1]), q(quals + seq_off[r], quals + seq_off[r + 1]);
    std::string gapped;
    int32_t pos = read_start[r];
    size_t k = 0;
    std::vector<CigarElement> cig;
    for (int c = cigar_off[r]; c < cigar_off[r + 1]; c++) {
      cig.push_back(CigarElement(cigar_type[c], cigar_len[c]));
      if (cigar_type[c] == 'D') { gapped.append(cigar_len[c], '-'); pos += cigar_len[c]; }
      else {
        gapped.append(seq, k, cigar_len[c]);
        k += cigar_len[c];
        if (cigar_type[c] != 'I') pos += cigar_len[c];
      }
    }
    Alignment a(read_start[r], read_stop ? read_stop[r] : pos - 1, rev_strand != NULL && rev_strand[r] != 0, "r" + std::to_string(name_id[r]), q, seq, gapped);
    a.set_cigar_list(cig);
    a.set_hap_gen_info(std::vector<bool>(1, use_for_haps == NULL || use_for_haps[r] != 0));
    alns.push_back(a);
    p1[sample_label[r]].push_back(log_p1[r]);
    p2[sample_label[r]].push_back(log_p2[r]);
  }
  Region region("chrS", region_start, region_stop, period, "STR");
  RegionGroup group(region);
  h->models.push_back(new StutterModel(stutter[0], stutter[1], stutter[2], stutter[3], stutter[4], stutter[5], period));
  if (ref_vcf_path != NULL) h->ref_vcf = new VCF::VCFReader(ref_vcf_path);
  h->g = new SeqStutterGenotyper(group, haploid != 0, reassemble_flanks != 0, alns, p1, p2, h->names, h->chrom_seq,
                                 h->models, h->ref_vcf, h->log);
  return h;
}

void ref_sg_destroy(void* hv) { delete static_cast<RefSG*>(hv); }
int32_t ref_sg_initialized(void* hv) { return static_cast<RefSG*>(hv)->g->initialized_ ? 1 : 0; }
int32_t ref_sg_num_blocks(void* hv) { return (int32_t)static_cast<RefSG*>(hv)->g->hap_blocks_.size(); }
int32_t ref_sg_num_haps(void* hv) { return static_cast<RefSG*>(hv)->g->num_alleles_; }
int32_t ref_sg_num_pools(void* hv) { return static_cast<RefSG*>(hv)->g->pooler_.num_pools(); }

/* info = {start, end, period (0 = flank block), n_options, total sequence bytes} */
void ref_sg_block_info(void* hv, int32_t b, int32_t* info) {
  HapBlock* blk = static_cast<RefSG*>(hv)->g->hap_blocks_[b];
  info[0] = blk->start();
  info[1] = blk->end();
  info[2] = blk->get_repeat_info() != NULL ? blk->get_repeat_info()->get_period() : 0;
  info[3] = blk->num_options();
  int32_t total = 0;
  for (int o = 0; o < blk->num_options(); o++) total += (int32_t)blk->get_seq(o).size();
  info[4] = total;
}
void ref_sg_block_seqs(void* hv, int32_t b, int32_t* off, char* seqs) {
  HapBlock* blk = static_cast<RefSG*>(hv)->g->hap_blocks_[b];
  off[0] = 0;
  for (int o = 0; o < blk->num_options(); o++) {
    const std::string& s = blk->get_seq(o);
    std::memcpy(seqs + off[o], s.data(), s.size());
    off[o + 1] = off[o] + (int32_t)s.size();
  }
}

int32_t ref_sg_genotype(void* hv, int32_t max_total_haps, int32_t max_flank_haps, double min_flank_freq) {
  RefSG* h = static_cast<RefSG*>(hv);
  return h->g->genotype(max_total_haps, max_flank_haps, min_flank_freq, h->log) ? 1 : 0;
}

int32_t ref_sg_recompute_stutter_models(void* hv, int32_t max_total_haps, int32_t max_flank_haps, double min_flank_freq,
                                        int32_t max_em_iter, double abs_ll, double frac_ll) {
  RefSG* h = static_cast<RefSG*>(hv);
  return h->g->recompute_stutter_models(h->log, max_total_haps, max_flank_haps, min_flank_freq, max_em_iter, abs_ll, frac_ll) ? 1 : 0;
}
/* the six parameters of the (first) repeat block's current stutter model */
void ref_sg_stutter_params(void* hv, double* out) {
  SeqStutterGenotyper* g = static_cast<RefSG*>(hv)->g;
  for (size_t b = 0; b < g->hap_blocks_.size(); b++)
    if (g->hap_blocks_[b]->get_repeat_info() != NULL) {
      StutterModel* m = g->hap_blocks_[b]->get_repeat_info()->get_stutter_model();
      out[0] = m->get_parameter(true, 'P'); out[1] = m->get_parameter(true, 'U'); out[2] = m->get_parameter(true, 'D');
      out[3] = m->get_parameter(false, 'P'); out[4] = m->get_parameter(false, 'U'); out[5] = m->get_parameter(false, 'D');
      return;
    }
}

/* Member arrays after genotype(): log_aln_probs_ [R*H], seed_positions_ [R], pool_index_ [R],
 * log_sample_posteriors_ [S*H*H], sample_total_LLs_ [S], optimal haplotypes [S*2], call_sample_ flags [S]. */
void ref_sg_results(void* hv, double* read_ll, int32_t* seeds, int32_t* pool_index, double* post, double* sample_ll,
                    int32_t* best, uint8_t* call_sample_ok) {
  SeqStutterGenotyper* g = static_cast<RefSG*>(hv)->g;
  const int R = g->num_reads_, H = g->num_alleles_, S = g->num_samples_;
  if (read_ll) std::memcpy(read_ll, g->log_aln_probs_, sizeof(double) * R * H);
  for (int r = 0; r < R; r++) {
    if (seeds) seeds[r] = g->seed_positions_[r];
    if (pool_index) pool_index[r] = g->pool_index_[r];
  }
  if (post) std::memcpy(post, g->log_sample_posteriors_, sizeof(double) * S * H * H);
  if (sample_ll) std::memcpy(sample_ll, g->sample_total_LLs_, sizeof(double) * S);
  if (best) {
    std::vector<std::pair<int, int> > gts;
    g->get_optimal_haplotypes(gts);
    for (int s = 0; s < S; s++) { best[2 * s] = gts[s].first; best[2 * s + 1] = gts[s].second; }
  }
  if (call_sample_ok)
    for (int s = 0; s < S; s++) call_sample_ok[s] = g->call_sample_[s].empty() ? 1 : 0;
}

/* write_vcf_record into a string: the record(s) are taken from the writer's reorder heap, so no
 * BGZF stream is ever opened.  Returns the number of bytes (excluding NUL), or -needed if cap is short. */
/* flags = the static Genotyper::OUTPUT_* switches: {GLS, PLS, PHASED_GLS, ALLREADS, MALLREADS, FILTERS, HAPLOTYPE_DATA} */
void ref_sg_set_output_flags(const int32_t* flags, double max_flank_indel_frac) {
  Genotyper::OUTPUT_GLS = flags[0]; Genotyper::OUTPUT_PLS = flags[1]; Genotyper::OUTPUT_PHASED_GLS = flags[2];
  Genotyper::OUTPUT_ALLREADS = flags[3]; Genotyper::OUTPUT_MALLREADS = flags[4]; Genotyper::OUTPUT_FILTERS = flags[5];
  Genotyper::OUTPUT_HAPLOTYPE_DATA = flags[6];
  Genotyper::MAX_FLANK_INDEL_FRAC = (float)max_flank_indel_frac;
}

int32_t ref_sg_write_vcf(void* hv, char* out, int32_t cap) {
  RefSG* h = static_cast<RefSG*>(hv);
  VCFWriter w;
  w.open_ = true;
  std::ostringstream html;
  h->g->write_vcf_record(h->names, h->chrom_seq, false, false, html, &w, h->log);
  std::string text;
  // records leave the heap in position order
  while (!w.record_heap_.empty()) {
    std::pop_heap(w.record_heap_.begin(), w.record_heap_.end(), tuple_comparator);
    RecordTuple* best = w.record_heap_.back();
    w.record_heap_.pop_back();
    text += best->text();
    text += "\n";
    delete best;
  }
  w.open_ = false;
  if ((int32_t)text.size() + 1 > cap) return -(int32_t)text.size() - 1;
  std::memcpy(out, text.c_str(), text.size() + 1);
  return (int32_t)text.size();
}

/* the third-party arithmetic behind AB / FS and the CIGAR window of ALLREADS, called directly */
double ref_allele_bias(int32_t a, int32_t b) {
  const int total = a + b;
  if (total == 0) return 1;
  if (a == b) return 0.0;
  return log10(std::min(1.0, 2 * bdtr(std::min(a, b), total, 0.5)));   // compute_allele_bias is private: same three lines
}
double ref_fisher_two_sided(int32_t n11, int32_t n12, int32_t n21, int32_t n22) {
  double left, right, two;
  kt_fisher_exact(n11, n12, n21, n22, &left, &right, &two);
  return two;
}
int32_t ref_extract_cigar(const char* type, const int32_t* len, int32_t n, int32_t cigar_start, int32_t region_start,
                          int32_t region_end, int32_t* bp_diff) {
  std::vector<CigarElement> cig;
  for (int i = 0; i < n; i++) cig.push_back(CigarElement(type[i], len[i]));
  int d = 0;
  const bool ok = ExtractCigar(cig, cigar_start, region_start, region_end, d);
  *bp_diff = d;
  return ok ? 1 : 0;
}

/* NeedlemanWunsch::Align itself on one pair: the alignment rows folded into one operation per column
 * ('I' where the reference row has a gap, 'D' where the read row has one, 'M' otherwise). */
int32_t ref_nw_align(const char* ref, int32_t L1, const char* read, int32_t L2, int32_t use_ref_end_penalty, char* ops, float* score) {
  std::string ref_al, read_al;
  std::vector<CigarOp> cigar;
  NeedlemanWunsch::Align(std::string(ref, ref + L1), std::string(read, read + L2), ref_al, read_al, score, cigar,
                         use_ref_end_penalty != 0);
  for (size_t i = 0; i < ref_al.size(); i++) ops[i] = ref_al[i] == '-' ? 'I' : (read_al[i] == '-' ? 'D' : 'M');
  ops[ref_al.size()] = 0;
  return (int32_t)ref_al.size();
}

/* One BAM-level alignment through the reference's left-alignment steps (genotyper_bam_processor.cpp:53-67):
 * TrimAlignment (bam_io.cpp:384-477) when do_trim, then convertAlignment for reads whose CIGAR is all M / = , else
 * realign (SeqAlignment/AlignmentOps.cpp:14-167).  The BamAlignment is assembled in memory.
 * Returns -1 if nothing is left after trimming, 0 if realign() failed, 1 = converted, 2 = realigned.
 * out_pos = {start, stop}; strings NUL-terminated; CIGAR in out_ctype / out_clen (*n_out_cigar runs). */
int32_t ref_left_align_one(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar,
                           const char* cigar_type, const int32_t* cigar_len, const char* chrom_seq, int32_t do_trim,
                           int32_t trim_start, int32_t trim_stop, int32_t* out_pos, char* out_seq, char* out_qual,
                           char* out_aln, int32_t* n_out_cigar, char* out_ctype, int32_t* out_clen) {
  BamAlignment b;
  b.bases_ = bases;
  b.qualities_ = quals;
  for (int i = 0; i < n_cigar; i++) b.cigar_ops_.push_back(CigarOp(cigar_type[i], cigar_len[i]));
  b.built_ = true;
  b.length_ = (int32_t)b.bases_.size();
  b.pos_ = pos;
  b.end_pos_ = end_pos;
  const char* qname = "read";
  b.b_->data = (uint8_t*)malloc(8);
  memcpy(b.b_->data, qname, 5);
  b.b_->l_data = 5; b.b_->m_data = 8;
  b.b_->core.l_qname = 5;
  if (do_trim) b.TrimAlignment(trim_start, trim_stop);
  if (b.Length() == 0) return -1;
  Alignment out("read");
  int32_t how;
  const std::string chrom(chrom_seq);
  if (b.MatchesReference()) { convertAlignment(b, chrom, out); how = 1; }
  else how = realign(b, chrom, out) ? 2 : 0;
  out_pos[0] = out.get_start();
  out_pos[1] = out.get_stop();
  strcpy(out_seq, out.get_sequence().c_str());
  strcpy(out_qual, out.get_base_qualities().c_str());
  strcpy(out_aln, out.get_alignment().c_str());
  *n_out_cigar = (int32_t)out.get_cigar_list().size();
  for (int i = 0; i < *n_out_cigar; i++) { out_ctype[i] = out.get_cigar_list()[i].get_type(); out_clen[i] = out.get_cigar_list()[i].get_num(); }
  return how;
}

/* The per-sample assembly step of assemble_flanks with the reference's own DebruijnGraph (same contract as
 * hipstr_flank_assemble). */
int32_t ref_flank_assemble(const char* ref_seq, int32_t n_seqs, const char* const* seqs, int32_t min_kmer, int32_t max_kmer,
                           int32_t* k_used, int32_t max_paths, int32_t path_cap, char* paths, int32_t* weights) {
  const std::string ref(ref_seq);
  const int max_k = std::min(max_kmer, ref.size() == 0 ? -1 : (int)ref.size() - 1);
  int kmer_length;
  if (!DebruijnGraph::calc_kmer_length(ref, min_kmer, max_k, kmer_length)) return -1;
  for (int k = kmer_length; k <= max_k; k++) {
    DebruijnGraph assembler(k, ref);
    for (int i = 0; i < n_seqs; i++) {
      std::string s(seqs[i]);
      if (!s.empty()) assembler.add_string(s);
    }
    assembler.prune_edges(0.02, 2);
    if (!assembler.has_cycles() && assembler.is_source_ok() && assembler.is_sink_ok()) {
      std::vector<std::pair<std::string, int> > found;
      assembler.enumerate_paths(2, max_paths, found);
      *k_used = k;
      for (size_t i = 0; i < found.size(); i++) {
        strcpy(paths + i * (size_t)path_cap, found[i].first.c_str());
        weights[i] = found[i].second;
      }
      return (int32_t)found.size();
    }
  }
  return -3;
}

/* The log the reference wrote for this locus (diagnostics in test failures). */
int32_t ref_sg_log(void* hv, char* out, int32_t cap) {
  std::string s = static_cast<RefSG*>(hv)->log.str();
  if ((int32_t)s.size() + 1 > cap) s = s.substr(s.size() + 1 - cap);
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}

}  // extern "C"
