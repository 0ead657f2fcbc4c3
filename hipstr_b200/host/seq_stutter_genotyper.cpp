/*
 * seq_stutter_genotyper.cpp -- see seq_stutter_genotyper.h.  Host control loop of seam B1 above the
 * C-ABI: every alignment, posterior and traceback is a batched device call; this file only holds the
 * reference's per-locus decisions, restated so that many loci advance together.
 */
#include "seq_stutter_genotyper.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <atomic>
#include <cstdlib>
#include <sstream>
#include <string_view>
#include <thread>
#include <emmintrin.h>

#include "../csrc/flatten.h"
#include "flank_assembler.h"
#include "haplotype_generator.h"

namespace hipstr {

namespace {

}  // namespace
int host_threads() { return host_thread_budget(); }
void parallel_for(size_t n, const std::function<void(size_t)>& fn) { parallel_run(n, host_threads(), fn); }
void GenotyperBatch::account_device_call() {
  int64_t h2d = 0, d2h = 0;
  int32_t launches = 0;
  hipstr_last_traffic(ctx_, &h2d, &d2h, &launches);
  h2d_bytes += h2d; d2h_bytes += d2h; gpu_launches += launches;
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
namespace {

bool order_by_length_and_sequence(const std::string& a, const std::string& b) {   // stringops.cpp:35-39
  if (a.size() != b.size()) return a.size() < b.size();
  return a.compare(b) < 0;
}

/* ---- Needleman-Wunsch of a haplotype against the reference haplotype -------------------------- */
const float kMatch = 2.0f, kMismatch = -2.0f, kGapOpen = 5.0f, kGapExtend = 0.125f, kLarge = 1000000.0f;

inline int base_code(char c) {   // NeedlemanWunsch.cpp:105-123 (anything else scores like N)
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}
inline float pair_score(int a, int b) { return (a == 4 || b == 4 || a == b) ? kMatch : kMismatch; }

/* The reference's three-way choice (NeedlemanWunsch.cpp:125-147): ties prefer the LATER matrix
 * between the second and third, the FIRST matrix against either. */
inline float pick3(float s1, float s2, float s3, int8_t* which) {
  if (s2 > s1) {
    if (s2 > s3) { *which = 1; return s2; }
    *which = 2;
    return s3;
  }
  if (s3 > s1) { *which = 2; return s3; }
  *which = 0;
  return s1;
}

/* Move indels the aligner left in the upstream flank to the right until they touch the repeat block
 * (Haplotype::adjust_indels, Haplotype.cpp:8-56). */
void shift_indels_toward_repeat(std::string& ref_row, std::string& alt_row, int32_t flank_start, int32_t repeat_start) {
  const size_t n = alt_row.size();
  int32_t ref_pos = flank_start;
  size_t col = 0;
  while (col < n) {
    if (alt_row[col] == '-' && ref_pos < repeat_start) {          // a deletion run
      size_t run_end = col;
      while (run_end < n && alt_row[run_end] == '-') run_end++;
      const int32_t run_len = (int32_t)(run_end - col);
      int32_t pos = ref_pos;
      size_t head = col;
      while (run_end < n && pos < repeat_start && ref_row[head] == ref_row[run_end]) {
        alt_row[head++] = alt_row[run_end];
        alt_row[run_end++] = '-';
        pos++;
      }
      col = run_end;
      ref_pos = pos + run_len;
    } else if (ref_row[col] == '-' && ref_pos < repeat_start) {   // an insertion run
      size_t run_end = col;
      while (run_end < n && ref_row[run_end] == '-') run_end++;
      int32_t pos = ref_pos;
      size_t head = col;
      while (run_end < n && pos < repeat_start && alt_row[head] == alt_row[run_end]) {
        ref_row[head++] = ref_row[run_end];
        ref_row[run_end++] = '-';
        pos++;
      }
      col = run_end;
      ref_pos = pos;
    } else {
      if (ref_row[col] != '-') ref_pos++;
      col++;
    }
  }
}

}  // namespace

/* ---- HapBlock ---------------------------------------------------------------------------------- */
bool HapBlock::contains(const std::string& s) const { return std::find(seqs.begin(), seqs.end(), s) != seqs.end(); }

HapBlock HapBlock::remove_alleles(const std::vector<int>& allele_indices) const {
  HapBlock out;
  out.start = start; out.end = end; out.period = period;
  std::memcpy(out.stutter, stutter, sizeof(stutter));
  for (size_t i = 0; i < seqs.size(); i++)
    if (i == 0 || std::find(allele_indices.begin(), allele_indices.end(), (int)i) == allele_indices.end()) out.seqs.push_back(seqs[i]);
  return out;
}

/* Four cells of the three-way choice above: the value, and which matrix it came from (0, 1, 2) as an integer vector. */
inline __m128 blend4(__m128 mask, __m128 a, __m128 b) { return _mm_or_ps(_mm_and_ps(mask, a), _mm_andnot_ps(mask, b)); }
inline __m128 pick3x4(__m128 s1, __m128 s2, __m128 s3, __m128i* which) {
  const __m128 second = _mm_cmpgt_ps(s2, s1);
  const __m128 a = blend4(second, s2, s1);
  const __m128 third = _mm_or_ps(_mm_cmpgt_ps(s3, a), _mm_and_ps(second, _mm_cmpeq_ps(s3, a)));   // ties: third over second, first over third
  const __m128i t = _mm_castps_si128(third);
  *which = _mm_or_si128(_mm_and_si128(t, _mm_set1_epi32(2)), _mm_andnot_si128(t, _mm_and_si128(_mm_castps_si128(second), _mm_set1_epi32(1))));
  return blend4(third, s3, a);
}

std::string hap_aln_to_ref(const std::string& ref_hap, const std::string& alt_hap, int32_t first_block_start,
                           int32_t repeat_block_start) {
  const int L1 = (int)ref_hap.size(), L2 = (int)alt_hap.size(), W = L1 + 1;
  const int P = ((W + 3) & ~3) + 8;   // pitch of the work rows: whole vectors plus slack for the shifted loads
  const size_t cells = (size_t)W * (L2 + 1);
  // M: bases paired; X: reference base against a gap; Y: alternate base against a gap.  Only two rows of scores are live;
  // the three predecessor choices of a cell share one byte (2 bits each: M, X, Y from the low bits up).
  //
  // A row is computed four cells at a time (SSE2, part of x86-64).  M and Y only look at the row above.  X looks at its
  // left neighbour, X[j] = max(M[j-1] - open, Y[j-1] - open, X[j-1] - extend), which in the frame X[j] + j * extend is a
  // running maximum of t[j] = max(M[j-1], Y[j-1]) - open + j * extend: one scalar max per cell.  Every score is a multiple
  // of 1/8 far below 2^21, so binary32 arithmetic is exact here and the re-association changes no value; the predecessor
  // choices, ties included, are then taken cell by cell from the finished values exactly as pick3 takes them.
  static thread_local std::vector<float> work;
  static thread_local std::vector<int32_t> choice;
  static thread_local std::vector<uint8_t> from;
  work.assign((size_t)P * 13, 0.0f);
  choice.assign((size_t)P * 2, 0);
  from.resize(cells + 16);
  float *pM = work.data(), *pX = pM + P, *pY = pX + P, *cM = pY + P, *cX = cM + P, *cY = cX + P;
  float *sub = cY + P;                      // [5][P]: pair score of alternate base code a against reference base j-1
  float *t = sub + 5 * (size_t)P, *ramp = t + P;   // ramp[j] = j * extend
  int32_t *wM = choice.data(), *wY = wM + P;
  pM[0] = 0.0f; pX[0] = -kLarge; pY[0] = -kLarge;
  for (int j = 1; j <= L1; j++) { pX[j] = -kGapOpen - (j - 1) * kGapExtend; pY[j] = -kLarge; pM[j] = -kLarge; }
  for (int j = 1; j <= L1; j++) {
    const int r = base_code(ref_hap[j - 1]);
    for (int a = 0; a < 5; a++) sub[(size_t)a * P + j] = pair_score(r, a);
  }
  for (int j = 0; j < P; j++) ramp[j] = j * kGapExtend;
  const __m128 open = _mm_set1_ps(kGapOpen), extend = _mm_set1_ps(kGapExtend);
  for (int i = 1; i <= L2; i++) {
    uint8_t* frow = from.data() + (size_t)i * W;
    cY[0] = -kGapOpen - (i - 1) * kGapExtend; cX[0] = -kLarge; cM[0] = -kLarge;
    const float* srow = sub + (size_t)base_code(alt_hap[i - 1]) * P;
    for (int j = 1; j <= L1; j += 4) {
      __m128i wm, wy;
      const __m128 m = pick3x4(_mm_loadu_ps(pM + j - 1), _mm_loadu_ps(pX + j - 1), _mm_loadu_ps(pY + j - 1), &wm);
      const __m128 y = pick3x4(_mm_sub_ps(_mm_loadu_ps(pM + j), open), _mm_sub_ps(_mm_loadu_ps(pX + j), open),
                               _mm_sub_ps(_mm_loadu_ps(pY + j), extend), &wy);
      _mm_storeu_ps(cM + j, _mm_add_ps(m, _mm_loadu_ps(srow + j)));
      _mm_storeu_ps(cY + j, y);
      _mm_storeu_si128(reinterpret_cast<__m128i*>(wM + j), wm);
      _mm_storeu_si128(reinterpret_cast<__m128i*>(wY + j), wy);
    }
    for (int j = 1; j <= L1; j += 4) {
      const __m128 s1 = _mm_sub_ps(_mm_loadu_ps(cM + j - 1), open), s3 = _mm_sub_ps(_mm_loadu_ps(cY + j - 1), open);
      _mm_storeu_ps(t + j, _mm_add_ps(_mm_max_ps(s1, s3), _mm_loadu_ps(ramp + j)));
    }
    float run = cX[0];
    for (int j = 1; j <= L1; j++) { run = run > t[j] ? run : t[j]; cX[j] = run - ramp[j]; }
    for (int j = 1; j <= L1; j += 4) {
      __m128i wx;
      (void)pick3x4(_mm_sub_ps(_mm_loadu_ps(cM + j - 1), open), _mm_sub_ps(_mm_loadu_ps(cX + j - 1), extend),
                    _mm_sub_ps(_mm_loadu_ps(cY + j - 1), open), &wx);
      const __m128i f = _mm_or_si128(_mm_or_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(wM + j)), _mm_slli_epi32(wx, 2)),
                                     _mm_slli_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(wY + j)), 4));
      const int32_t four = _mm_cvtsi128_si32(_mm_packus_epi16(_mm_packs_epi32(f, f), _mm_setzero_si128()));
      std::memcpy(frow + j, &four, 4);   // may spill up to 3 bytes into the next row, which is written after this one
    }
    frow[0] = (uint8_t)(2 << 4);
    std::swap(pM, cM); std::swap(pX, cX); std::swap(pY, cY);
  }
  // end-to-end alignment: stop in the corner (findOptimalStopEndPenalty)
  int kind = 0;
  float best = pM[L1];
  if (pX[L1] > best) { best = pX[L1]; kind = 1; }
  if (pY[L1] > best) { best = pY[L1]; kind = 2; }
  std::string ref_row, alt_row;   // built back to front
  ref_row.reserve((size_t)L1 + L2); alt_row.reserve((size_t)L1 + L2);
  int row = L2, col = L1;
  while (row > 0) {
    const uint8_t f = from[(size_t)row * W + col];
    if (kind == 0 && col > 0) { ref_row += ref_hap[col - 1]; alt_row += alt_hap[row - 1]; kind = f & 3; row--; col--; }
    else if (kind == 1 && col > 0) { ref_row += ref_hap[col - 1]; alt_row += '-'; kind = (f >> 2) & 3; col--; }
    else if (kind == 2) { ref_row += '-'; alt_row += alt_hap[row - 1]; kind = (f >> 4) & 3; row--; }
    else return std::string();   // the reference dies here ("Invalid matrix type")
  }
  for (; col > 0; col--) { ref_row += ref_hap[col - 1]; alt_row += '-'; }
  std::reverse(ref_row.begin(), ref_row.end());
  std::reverse(alt_row.begin(), alt_row.end());
  shift_indels_toward_repeat(ref_row, alt_row, first_block_start, repeat_block_start);
  std::string info(alt_row.size(), 'M');
  for (size_t i = 0; i < alt_row.size(); i++)
    if (ref_row[i] == '-') info[i] = 'I';
    else if (alt_row[i] == '-') info[i] = 'D';
  return info;
}

/* ---- one locus ---------------------------------------------------------------------------------- */
std::string SeqStutterGenotyper::hap_seq(int hap) const {
  std::vector<int32_t> n(hap_blocks_.size()), opt(hap_blocks_.size());
  for (size_t b = 0; b < hap_blocks_.size(); b++) n[b] = hap_blocks_[b].num_options();
  haplotype_options((int)n.size(), n.data(), hap, opt.data());
  std::string s;
  for (size_t b = 0; b < hap_blocks_.size(); b++) s += hap_blocks_[b].seqs[opt[b]];
  return s;
}

void SeqStutterGenotyper::haps_to_alleles(int block_index, std::vector<int>& allele_indices) const {
  std::vector<int32_t> n(hap_blocks_.size()), opt(hap_blocks_.size());
  for (size_t b = 0; b < hap_blocks_.size(); b++) n[b] = hap_blocks_[b].num_options();
  allele_indices.resize(num_alleles_);
  for (int h = 0; h < num_alleles_; h++) {
    haplotype_options((int)n.size(), n.data(), h, opt.data());
    allele_indices[h] = opt[block_index];
  }
}

void SeqStutterGenotyper::rebuild_hap_aln_info(const std::map<std::string, std::string>* known) {
  hap_aln_info_.assign(num_alleles_, std::string());
  const std::string ref = hap_seq(0);
  // Haplotype::adjust_indels only knows flank / repeat / flank haplotypes (Haplotype.cpp:9 asserts 3 blocks)
  const int32_t first = hap_blocks_.front().start;
  const int32_t rep = hap_blocks_.size() > 1 ? hap_blocks_[1].start : hap_blocks_.front().end;
  for (int h = 0; h < num_alleles_; h++) {
    const std::string seq = hap_seq(h);
    if (known) {
      auto it = known->find(seq);
      if (it != known->end()) { hap_aln_info_[h] = it->second; continue; }
    }
    hap_aln_info_[h] = hap_aln_to_ref(ref, seq, first, rep);
  }
  hap_aln_index_.assign(num_alleles_, std::vector<int32_t>());
  for (int h = 0; h < num_alleles_; h++) {
    hap_aln_index_[h].resize(3 * (hap_aln_info_[h].size() + 1));
    if (hipstr_hap_aln_index(hap_aln_info_[h].c_str(), (int32_t)hap_aln_info_[h].size(), hap_aln_index_[h].data()) != HIPSTR_OK)
      hap_aln_index_[h].clear();   // an empty string ("Invalid matrix type") or one with other operations: the stepping form
  }
}

int SeqStutterGenotyper::best_hap_of_read(int r) const {
  const double log_one_half = host_tables().log_one_half;
  const int s = sample_label_[r], hap_a = optimal_haps_[2 * s], hap_b = optimal_haps_[2 * s + 1];
  const double* ll = &log_aln_probs_[(size_t)r * num_alleles_];
  return (log_one_half + log_p1_[r] + ll[hap_a] > log_one_half + log_p2_[r] + ll[hap_b]) ? hap_a : hap_b;
}

bool SeqStutterGenotyper::collect_missing_traces() {
  missing_traces_.clear();
  missing_trace_read_.clear();
  std::unordered_set<uint64_t> wanted;
  for (int r = 0; r < num_reads_; r++) {
    if (seed_positions_[r] < 0) continue;
    std::pair<int, int> key(pool_index_[r], best_hap_of_read(r));
    if (trace_cache_.count(key) == 0 && wanted.insert(((uint64_t)(uint32_t)key.first << 32) | (uint32_t)key.second).second)
      missing_traces_.push_back(key);
  }
  return missing_traces_.empty();
}

void SeqStutterGenotyper::get_stutter_candidate_alleles(int block_index, std::vector<std::string>& candidate_seqs) {
  const HapBlock& block = hap_blocks_[block_index];
  std::vector<int> sample_counts(num_samples_, 0);
  std::vector<std::map<std::string, int> > sample_stutter_counts(num_samples_);
  for (int r = 0; r < num_reads_; r++) {
    if (seed_positions_[r] < 0) continue;
    const AlignmentTrace& trace = trace_cache_.at(std::make_pair(pool_index_[r], best_hap_of_read(r)));
    if (trace.start < block.start && trace.stop > block.end) {
      if (trace.stutter_size[block_index] != 0) sample_stutter_counts[sample_label_[r]][std::string(trace.str_seq(block_index))]++;
      sample_counts[sample_label_[r]]++;
    }
  }
  std::set<std::string> candidate_set;   // artifacts seen at least twice and in >= 15 % of a sample's spanning reads
  for (int s = 0; s < num_samples_; s++)
    for (const auto& kv : sample_stutter_counts[s])
      if (kv.second >= 2 && 1.0 * kv.second / sample_counts[s] >= 0.15 && !block.contains(kv.first)) candidate_set.insert(kv.first);
  candidate_seqs.assign(candidate_set.begin(), candidate_set.end());
  if (!candidate_seqs.empty()) {
    std::ostringstream msg;
    msg << "Identified " << candidate_seqs.size() << " additional candidate alleles from stutter artifacts\n";
    for (const auto& s : candidate_seqs) msg << "\t" << s << "\n";
    log_ += msg.str();
  }
}

void SeqStutterGenotyper::get_unused_alleles(bool check_spanned, bool check_called,
                                             std::vector<std::vector<int> >& allele_indices, int& num_aff_blocks,
                                             int& num_aff_alleles) {
  allele_indices.clear();
  num_aff_blocks = num_aff_alleles = 0;
  std::vector<bool> aligned_read(num_samples_, false);
  for (int r = 0; r < num_reads_; r++)
    if (seed_positions_[r] >= 0) aligned_read[sample_label_[r]] = true;
  const double tolerance = 1e-10;   // mathops.cpp:10
  for (int b = 0; b < (int)hap_blocks_.size(); b++) {
    allele_indices.push_back(std::vector<int>());
    const HapBlock& block = hap_blocks_[b];
    if (block.num_options() == 1) continue;
    std::vector<int> hap_to_allele;
    haps_to_alleles(b, hap_to_allele);
    std::vector<bool> spanned(block.num_options(), false), called(block.num_options(), false);
    if (check_spanned) {   // alleles supported by a spanning read whose best alignment carries no stutter
      for (int r = 0; r < num_reads_; r++) {
        if (seed_positions_[r] < 0) continue;
        const AlignmentTrace& trace = trace_cache_.at(std::make_pair(pool_index_[r], best_hap_of_read(r)));
        if (!(trace.start < block.start && trace.stop > block.end)) continue;
        if (trace.stutter_size[b] != 0) continue;
        const int s = sample_label_[r], hap_a = optimal_haps_[2 * s], hap_b = optimal_haps_[2 * s + 1];
        int best_hap = hap_a;
        if (!haploid_ && hap_a != hap_b) {
          const double* ll = &log_aln_probs_[(size_t)r * num_alleles_];
          const double v1 = log_p1_[r] + ll[hap_a], v2 = log_p2_[r] + ll[hap_b];
          if (std::fabs(v1 - v2) > tolerance) best_hap = v1 > v2 ? hap_a : hap_b;
        }
        spanned[hap_to_allele[best_hap]] = true;
      }
    }
    if (check_called)
      for (int s = 0; s < num_samples_; s++)
        if (aligned_read[s] && call_sample_[s].empty()) {
          called[hap_to_allele[optimal_haps_[2 * s]]] = true;
          called[hap_to_allele[optimal_haps_[2 * s + 1]]] = true;
        }
    bool affected = false;
    for (int a = 1; a < block.num_options(); a++)
      if ((check_spanned && !spanned[a]) || (check_called && !called[a])) {
        allele_indices.back().push_back(a);
        affected = true;
        num_aff_alleles++;
      }
    if (affected) num_aff_blocks++;
  }
}

bool SeqStutterGenotyper::add_and_remove_alleles(const std::vector<std::vector<int> >& alleles_to_remove,
                                                 const std::vector<std::vector<std::string> >& alleles_to_add,
                                                 const std::vector<uint8_t>* realign_pool, const std::vector<uint8_t>* copy_read) {
  const int old_H = num_alleles_;
  // haplotypes are matched across the change by SEQUENCE; of two old haplotypes with one sequence the later wins
  std::map<std::string, int> old_index;
  std::map<std::string, std::string> old_info;
  for (int h = 0; h < old_H; h++) {
    const std::string seq = hap_seq(h);
    old_index[seq] = h;
    old_info[seq] = hap_aln_info_[h];
  }
  bool added_seq = false;
  for (size_t b = 0; b < hap_blocks_.size(); b++) {
    hap_blocks_[b] = hap_blocks_[b].remove_alleles(alleles_to_remove[b]);
    for (const std::string& s : alleles_to_add[b]) { hap_blocks_[b].seqs.push_back(s); added_seq = true; }
  }
  int new_H = 1;
  for (const HapBlock& b : hap_blocks_) new_H *= b.num_options();
  num_alleles_ = new_H;
  std::vector<int> allele_mapping(old_H, -1);
  realign_hap_.assign(new_H, 0);
  for (int h = 0; h < new_H; h++) {
    auto match = old_index.find(hap_seq(h));
    if (match == old_index.end()) realign_hap_[h] = 1;
    else allele_mapping[match->second] = h;
  }
  // surviving columns keep their likelihoods; new columns start at -100000 (seq_stutter_genotyper.cpp:374)
  std::vector<double> fixed((size_t)num_reads_ * new_H, -100000.0);
  for (int r = 0; r < num_reads_; r++)
    for (int j = 0; j < old_H; j++)
      if (allele_mapping[j] != -1) fixed[(size_t)r * new_H + allele_mapping[j]] = log_aln_probs_[(size_t)r * old_H + j];
  log_aln_probs_.swap(fixed);
  TraceCache remapped;
  remapped.reserve(trace_cache_.size());
  for (auto& kv : trace_cache_) {
    const int h = allele_mapping[kv.first.second];
    if (h != -1) remapped[std::make_pair(kv.first.first, h)] = std::move(kv.second);
  }
  trace_cache_.swap(remapped);
  log_sample_posteriors_.assign((size_t)num_samples_ * new_H * new_H, 0.0);
  rebuild_hap_aln_info(&old_info);
  if (realign_pool) realign_pool_ = *realign_pool; else realign_pool_.clear();
  if (copy_read) copy_read_ = *copy_read; else copy_read_.clear();
  return added_seq;
}

SeqStutterGenotyper::Request SeqStutterGenotyper::advance() {
  struct PhaseTimer {   // charges the time until the next phase change / return to the phase that was running
    SeqStutterGenotyper* g; Phase p; double t0;
    explicit PhaseTimer(SeqStutterGenotyper* g_) : g(g_), p(g_->phase_), t0(now_s()) {}
    ~PhaseTimer() { g->phase_seconds_[p] += now_s() - t0; }
  };
  for (;;) {
    PhaseTimer timer(this);
    switch (phase_) {
      case ALIGN_ALL:
        realign_hap_.assign(num_alleles_, 1);
        realign_pool_.clear();
        copy_read_.clear();
        log_aln_probs_.assign((size_t)num_reads_ * num_alleles_, 0.0);
        log_sample_posteriors_.assign((size_t)num_samples_ * num_alleles_ * num_alleles_, 0.0);
        log_ += "Aligning reads to each candidate haplotype\n";
        // with a reference panel the allele set is fixed: no stutter-allele discovery, no pruning (seq_stutter_genotyper.cpp:641-665)
        phase_ = fixed_alleles_ ? (reassemble_flanks_ ? ASSEMBLE_FLANKS : DONE) : STUTTER_ALLELES;
        rounds_++;
        return NEED_ALIGNMENT;
      case STUTTER_ALLELES: {   // id_and_align_to_stutter_alleles, seq_stutter_genotyper.cpp:570-601
        if (!collect_missing_traces()) return NEED_TRACES;
        std::vector<std::vector<int> > none(hap_blocks_.size());
        std::vector<std::vector<std::string> > stutter_seqs(hap_blocks_.size());
        int new_total_haps = num_alleles_;
        bool added = false;
        for (size_t b = 0; b < hap_blocks_.size(); b++) {
          if (hap_blocks_[b].period <= 0) continue;
          get_stutter_candidate_alleles((int)b, stutter_seqs[b]);
          added |= !stutter_seqs[b].empty();
          std::sort(stutter_seqs[b].begin(), stutter_seqs[b].end(), order_by_length_and_sequence);
          new_total_haps /= hap_blocks_[b].num_options();
          new_total_haps *= hap_blocks_[b].num_options() + (int)stutter_seqs[b].size();
        }
        if (!added) { phase_ = PRUNE_UNCALLED; break; }
        if (new_total_haps > max_total_haplotypes_) {
          std::ostringstream msg;
          msg << "Aborting genotyping of the locus as too many candidate haplotypes were found (# Found = " << new_total_haps
              << ", MAX = " << max_total_haplotypes_ << ")\n";
          log_ += msg.str();
          phase_ = FAILED;
          return NONE;
        }
        add_and_remove_alleles(none, stutter_seqs);
        rounds_++;
        return NEED_ALIGNMENT;
      }
      case PRUNE_UNCALLED: {    // seq_stutter_genotyper.cpp:646-654
        std::vector<std::vector<int> > unused;
        int blocks = 0, alleles = 0;
        get_unused_alleles(false, true, unused, blocks, alleles);
        phase_ = PRUNE_UNSPANNED;
        if (alleles == 0) break;
        std::ostringstream msg;
        msg << "Recomputing sample posteriors after removing " << alleles << " uncalled alleles across " << blocks << " blocks\n";
        log_ += msg.str();
        add_and_remove_alleles(unused, std::vector<std::vector<std::string> >(hap_blocks_.size()));
        return NEED_POSTERIORS;
      }
      case PRUNE_UNSPANNED: {   // seq_stutter_genotyper.cpp:656-664
        if (!collect_missing_traces()) return NEED_TRACES;
        std::vector<std::vector<int> > unused;
        int blocks = 0, alleles = 0;
        get_unused_alleles(true, false, unused, blocks, alleles);
        phase_ = reassemble_flanks_ ? ASSEMBLE_FLANKS : DONE;
        if (alleles == 0) break;
        std::ostringstream msg;
        msg << "Recomputing sample posteriors after removing " << alleles << " alleles with no spanning reads across " << blocks
            << " blocks\n";
        log_ += msg.str();
        add_and_remove_alleles(unused, std::vector<std::vector<std::string> >(hap_blocks_.size()));
        return NEED_POSTERIORS;
      }
      case ASSEMBLE_FLANKS: {   // assemble_flanks, seq_stutter_genotyper.cpp:40-217
        if (!collect_missing_traces()) return NEED_TRACES;
        const int outcome = assemble_flanks();
        if (outcome < 0) { phase_ = FAILED; return NONE; }
        if (outcome == 0) { phase_ = DONE; break; }
        phase_ = fixed_alleles_ ? DONE : ASSEMBLE_PRUNE;   // (:204: uncalled alleles are only removed without a reference panel)
        rounds_++;
        return NEED_ALIGNMENT;
      }
      case ASSEMBLE_PRUNE: {    // seq_stutter_genotyper.cpp:203-213
        std::vector<std::vector<int> > unused;
        int blocks = 0, alleles = 0;
        get_unused_alleles(false, true, unused, blocks, alleles);
        phase_ = DONE;
        if (alleles == 0) break;
        std::ostringstream msg;
        msg << "Recomputing sample posteriors after removing " << alleles << " uncalled alleles across " << blocks << " blocks\n";
        log_ += msg.str();
        add_and_remove_alleles(unused, std::vector<std::vector<std::string> >(hap_blocks_.size()));
        return NEED_POSTERIORS;
      }
      case DONE:
      case FAILED:
        return NONE;
    }
  }
}

int SeqStutterGenotyper::assemble_flanks() {
  const int kMinPathWeight = 2, kMinKmer = 10, kMaxKmer = 15;   // seq_stutter_genotyper.h:152-154
  log_ += "Reassembling flanking sequences\n";
  std::vector<std::vector<std::string> > alleles_to_add(hap_blocks_.size());
  std::vector<bool> realign_sample(num_samples_, false);
  int new_total_haps = num_alleles_;
  // reads are sample-major: [first_read[s], first_read[s+1]) belong to sample s
  std::vector<int> first_read(num_samples_ + 1, num_reads_);
  for (int r = num_reads_ - 1; r >= 0; r--) first_read[sample_label_[r]] = r;
  for (int s = num_samples_ - 1; s >= 0; s--) first_read[s] = std::min(first_read[s], first_read[s + 1]);
  // the trace of every read against its best haplotype, looked up once for both flanks
  std::vector<int32_t> read_trace(num_reads_, -1);   // slot in trace_cache_: reads of one pool share their trace
  for (int r = 0; r < num_reads_; r++)
    if (seed_positions_[r] >= 0) read_trace[r] = trace_cache_.slot_of(std::make_pair(pool_index_[r], best_hap_of_read(r)));
  for (int flank = 0; flank < 2; flank++) {
    const int block_index = flank == 0 ? 0 : (int)hap_blocks_.size() - 1;
    const std::string& ref_seq = hap_blocks_[block_index].seqs[0];
    const int max_k = std::min(kMaxKmer, ref_seq.empty() ? -1 : (int)ref_seq.size() - 1);
    new_total_haps /= hap_blocks_[block_index].num_options();
    int kmer_length;
    if (!FlankAssembler::calc_kmer_length(ref_seq, kMinKmer, max_k, kmer_length)) return -1;

    // Per DISTINCT flank sequence of the locus (the same few strings come back in every sample): is it a substring of the
    // reference flank, and which of its (k+1)-mers at k = kmer_length are not reference edges.  Every read is looked up
    // once; the per-sample work below only touches these records.
    struct FlankInfo {
      std::string_view seq;
      bool in_reference = false;
      std::vector<int> nonref_edges;   // ids (below) of its (k+1)-mers that are not reference edges, one per occurrence
      int stamp = -1, count = 0;   // sample that saw it last, reads of that sample carrying it
    };
    std::unordered_map<std::string_view, FlankInfo> flank_info;
    std::unordered_map<std::string_view, int> edge_id;   // the locus' distinct non-reference (k+1)-mers
    std::unordered_set<std::string_view> ref_edges;      // the reference flank's (k+1)-mers
    for (size_t c = 0; c + kmer_length + 1 <= ref_seq.size(); c++) ref_edges.insert(std::string_view(ref_seq).substr(c, kmer_length + 1));
    std::vector<FlankInfo*> read_info(num_reads_, nullptr), trace_info(trace_cache_.size(), nullptr);
    std::vector<uint8_t> trace_seen(trace_cache_.size(), 0);
    for (int r = 0; r < num_reads_; r++) {
      if (read_trace[r] < 0) continue;
      if (trace_seen[read_trace[r]]) { read_info[r] = trace_info[read_trace[r]]; continue; }
      trace_seen[read_trace[r]] = 1;
      const std::string_view sv = trace_cache_.in_slot(read_trace[r]).flank_seq(block_index);
      if (sv.empty()) continue;
      auto it = flank_info.find(sv);
      if (it == flank_info.end()) {
        FlankInfo fi;
        fi.seq = sv;
        fi.in_reference = ref_seq.find(sv) != std::string::npos;
        if (!fi.in_reference)
          for (size_t c = 0; c + kmer_length + 1 <= sv.size(); c++) {
            const std::string_view e = sv.substr(c, kmer_length + 1);
            if (!ref_edges.count(e)) fi.nonref_edges.push_back(edge_id.emplace(e, (int)edge_id.size()).first->second);
          }
        it = flank_info.emplace(sv, std::move(fi)).first;
      }
      read_info[r] = trace_info[read_trace[r]] = &it->second;
    }
    std::map<std::string, int> haplotype_indexes;           // alternate flank -> index
    std::vector<std::vector<int> > haplotype_to_sample;     // samples supporting each alternate flank
    std::vector<std::pair<std::string, int> > assembly_data;
    std::vector<FlankInfo*> flank_seqs;                     // the sample's distinct flank sequences in read order
    std::vector<int> edge_weight(edge_id.size(), 0), edge_stamp(edge_id.size(), -1);
    for (int s = 0; s < num_samples_; s++) {
      if (!call_sample_[s].empty()) continue;
      assembly_data.clear();
      bool acyclic = false;
      // the sample's flank sequences in read order, identical ones counted once (same graph, fewer k-mer walks)
      flank_seqs.clear();
      bool only_reference = true;
      for (int r = first_read[s]; r < first_read[s + 1]; r++) {
        FlankInfo* fi = read_info[r];
        if (!fi) continue;
        if (fi->stamp != s) { fi->stamp = s; fi->count = 0; flank_seqs.push_back(fi); only_reference &= fi->in_reference; }
        fi->count++;
      }
      // Reads that only repeat (part of) the reference flank add weight to reference edges and nothing else: the graph
      // is the reference path, acyclic at kmer_length by construction, with exactly one source-to-sink path -- the
      // outcome "no alternate flank" is known without building it.
      if (only_reference) continue;
      {
        // Same outcome, one step further: an edge of the graph is a (k+1)-mer; edges of the reference are never pruned
        // and every other edge is pruned below weight max(2, ceil(0.02 * strings)) (prune_edges, debruijn_graph.cpp:47-60).
        // If no non-reference (k+1)-mer of the sample's reads reaches that weight at k = kmer_length, pruning leaves the
        // bare reference path -- acyclic at that k, one source-to-sink path -- and the loop below would stop at its
        // first k with nothing to report.  Checking that needs a tally over the interned (k+1)-mers, not a graph.
        const int k = kmer_length;
        int num_strings = 1, heaviest = 0;
        for (const FlankInfo* fi : flank_seqs) {
          if ((int)fi->seq.size() <= k) continue;
          num_strings += fi->count;
          for (const int e : fi->nonref_edges) {
            if (edge_stamp[e] != s) { edge_stamp[e] = s; edge_weight[e] = 0; }
            edge_weight[e] += fi->count;
            heaviest = std::max(heaviest, edge_weight[e]);   // weights only grow: the last maximum is the final one
          }
        }
        const int min_weight = std::max(2, (int)std::ceil(0.02 * num_strings));
        const bool survives = heaviest >= min_weight;
        if (!survives) continue;
      }
      for (int k = kmer_length; k <= max_k; k++) {
        FlankAssembler assembler(k, ref_seq);
        for (const FlankInfo* fi : flank_seqs) assembler.add_string(fi->seq, 1, fi->count);
        assembler.prune_edges(0.02, 2);
        if (!assembler.has_cycles() && assembler.is_source_ok() && assembler.is_sink_ok()) {
          acyclic = true;
          assembler.enumerate_paths(kMinPathWeight, 10, assembly_data);
          break;
        }
      }
      if (!acyclic) { call_sample_[s] = "FLANK_ASSEMBLY_CYCLIC"; continue; }
      if (assembly_data.size() <= 1) continue;
      int total_depth = 0;
      for (const auto& path : assembly_data) total_depth += path.second;
      for (const auto& path : assembly_data) {
        if (path.first == ref_seq || !(path.second * 1.0 / total_depth > 0.25)) continue;
        if (ref_seq.size() != path.first.size()) {
          // a flank with an indel would clobber indels in the repeat: the sample is not genotyped
          call_sample_[s] = "FLANK_ASSEMBLY_INDEL";
          realign_sample[s] = false;
        } else {
          if (haplotype_indexes.find(path.first) == haplotype_indexes.end()) {
            const int index = (int)haplotype_indexes.size();
            haplotype_indexes[path.first] = index;
            haplotype_to_sample.push_back(std::vector<int>());
          }
          realign_sample[s] = true;
          haplotype_to_sample[haplotype_indexes[path.first]].push_back(s);
        }
      }
    }
    // flanks seen in too few samples are dropped and their samples are not genotyped
    for (auto it = haplotype_indexes.begin(); it != haplotype_indexes.end();) {
      const std::vector<int>& hap_samples = haplotype_to_sample[it->second];
      if (hap_samples.size() < min_flank_freq_ * num_samples_) {
        for (int s : hap_samples)
          if (call_sample_[s].empty()) { call_sample_[s] = "LOW_FREQUENCY_ALT_FLANK"; realign_sample[s] = false; }
        log_ += std::string("\tPruning low frequency ") + (flank == 0 ? "left" : "right") + " flank\t" + it->first + "\n";
        haplotype_indexes.erase(it++);
      } else
        ++it;
    }
    if (!haplotype_indexes.empty()) {
      if ((int)haplotype_indexes.size() > max_flank_haplotypes_) {
        log_ += std::string("Skipping locus with too many ") + (flank == 0 ? "left" : "right") + " alternate flanking sequences\n";
        return -1;
      }
      std::ostringstream msg;
      msg << "Identified " << haplotype_indexes.size() << " new " << (flank == 0 ? "left" : "right") << " flank haplotype(s)\n";
      for (const auto& kv : haplotype_indexes) {
        msg << "\t" << kv.first << "\t" << haplotype_to_sample[kv.second].size() << "\n";
        alleles_to_add[block_index].push_back(kv.first);
      }
      log_ += msg.str();
      new_total_haps *= 1 + (int)haplotype_indexes.size();
    }
  }
  if (new_total_haps > max_total_haplotypes_) {
    log_ += "Aborting genotyping of the locus as too many candidate haplotypes were found\n";
    return -1;
  }
  // realign a pool when any of its reads belongs to a sample with a new flank; update a read's
  // likelihoods only when its whole sample is realigned
  std::vector<uint8_t> realign_pools(num_pools_, 0), copy_reads(num_reads_, 0);
  for (int r = 0; r < num_reads_; r++) {
    const bool flag = realign_sample[sample_label_[r]];
    if (flag) realign_pools[pool_index_[r]] = 1;
    copy_reads[r] = flag ? 1 : 0;
  }
  const int realign_count = (int)std::count(realign_pools.begin(), realign_pools.end(), 1);
  if (realign_count == 0) return 0;
  std::ostringstream msg;
  msg << "Realigning " << realign_count << " out of " << num_pools_ << " read pools to polish flanking sequences\n";
  log_ += msg.str();
  add_and_remove_alleles(std::vector<std::vector<int> >(hap_blocks_.size()), alleles_to_add, &realign_pools, &copy_reads);
  return 1;
}

/* ---- batching ------------------------------------------------------------------------------------ */
namespace {

/* A hipstr_align_batch_t assembled from a list of loci. */
struct PackedBatch {
  std::vector<int32_t> locus_block_off{0}, locus_pool_off{0}, block_period, block_opt_off{0}, opt_seq_off{0}, pool_seq_off{0},
      pool_seed, block_start;
  std::vector<int64_t> locus_hap_off{0}, locus_out_off{0};
  std::vector<double> block_stutter;
  std::string opt_seq, pool_bases, pool_quals;
  std::vector<uint8_t> realign_pool, realign_hap;
  bool pool_masked = false, hap_masked = false;

  /* own_quality_reads: reads appended as extra pools that carry their own qualities (and their pool's seed) */
  void add(const SeqStutterGenotyper& g, const std::vector<uint8_t>* hap_mask, const std::vector<uint8_t>* pool_mask,
           const std::vector<int>* own_quality_reads = nullptr) {
    for (const HapBlock& b : g.hap_blocks_) {
      block_period.push_back(b.period);
      block_start.push_back(b.start);
      block_stutter.insert(block_stutter.end(), b.stutter, b.stutter + 6);
      for (const std::string& s : b.seqs) { opt_seq += s; opt_seq_off.push_back((int32_t)opt_seq.size()); }
      block_opt_off.push_back((int32_t)opt_seq_off.size() - 1);
    }
    locus_block_off.push_back((int32_t)block_period.size());
    const int32_t base = (int32_t)pool_bases.size();
    pool_bases += g.pool_bases_;
    pool_quals += g.pool_quals_;
    for (int p = 0; p < g.num_pools_; p++) {
      pool_seq_off.push_back(base + g.pool_seq_off_[p + 1]);
      pool_seed.push_back(g.pool_seed_[p]);
      const uint8_t m = (pool_mask && !pool_mask->empty()) ? (*pool_mask)[p] : 1;
      realign_pool.push_back(m);
      pool_masked |= (m == 0);
    }
    int extra = 0;
    if (own_quality_reads)
      for (int r : *own_quality_reads) {
        const int p = g.pool_index_[r];
        pool_bases.append(g.pool_bases_, g.pool_seq_off_[p], g.pool_seq_off_[p + 1] - g.pool_seq_off_[p]);
        pool_quals.append(g.read_quals_, g.read_seq_off_[r], g.read_seq_off_[r + 1] - g.read_seq_off_[r]);
        pool_seq_off.push_back((int32_t)pool_bases.size());
        pool_seed.push_back(g.pool_seed_[p]);
        realign_pool.push_back(1);
        extra++;
      }
    locus_pool_off.push_back((int32_t)pool_seed.size());
    for (int h = 0; h < g.num_alleles_; h++) {
      const uint8_t m = (hap_mask && !hap_mask->empty()) ? (*hap_mask)[h] : 1;
      realign_hap.push_back(m);
      hap_masked |= (m == 0);
    }
    locus_hap_off.push_back(locus_hap_off.back() + g.num_alleles_);
    locus_out_off.push_back(locus_out_off.back() + (int64_t)(g.num_pools_ + extra) * g.num_alleles_);
  }

  hipstr_align_batch_t view() const {
    hipstr_align_batch_t b;
    std::memset(&b, 0, sizeof(b));
    b.n_loci = (int32_t)locus_block_off.size() - 1;
    b.n_blocks = (int32_t)block_period.size();
    b.n_options = (int32_t)opt_seq_off.size() - 1;
    b.n_pools = (int32_t)pool_seed.size();
    b.n_haps = locus_hap_off.back();
    b.locus_block_off = locus_block_off.data();
    b.locus_pool_off = locus_pool_off.data();
    b.locus_hap_off = locus_hap_off.data();
    b.locus_out_off = locus_out_off.data();
    b.block_period = block_period.data();
    b.block_opt_off = block_opt_off.data();
    b.block_stutter = block_stutter.data();
    b.opt_seq_off = opt_seq_off.data();
    b.opt_seq = opt_seq.data();
    b.pool_seq_off = pool_seq_off.data();
    b.pool_bases = pool_bases.data();
    b.pool_quals = pool_quals.data();
    b.pool_seed = pool_seed.data();
    b.realign_pool = pool_masked ? realign_pool.data() : nullptr;
    b.realign_hap = hap_masked ? realign_hap.data() : nullptr;
    return b;
  }
};

}  // namespace

hipstr_status_t GenotyperBatch::init_reads(SeqStutterGenotyper& g, const hipstr_locus_reads_t* rd, int l, std::string& err) {
  {
    const int r0 = rd->locus_read_off[l], r1 = rd->locus_read_off[l + 1], R = r1 - r0;
    g.haploid_ = rd->haploid && rd->haploid[l];
    g.num_samples_ = rd->locus_sample_off[l + 1] - rd->locus_sample_off[l];
    g.num_reads_ = R;
    g.call_sample_.assign(g.num_samples_, std::string());
    g.sample_label_.assign(rd->sample_label + r0, rd->sample_label + r1);
    g.log_p1_.assign(rd->log_p1 + r0, rd->log_p1 + r1);
    g.log_p2_.assign(rd->log_p2 + r0, rd->log_p2 + r1);
    g.read_start_.assign(rd->read_start + r0, rd->read_start + r1);
    if (rd->rev_strand) g.rev_strand_.assign(rd->rev_strand + r0, rd->rev_strand + r1);
    else g.rev_strand_.assign(R, 0);
    // init(): a read is a second mate when it carries the name of the read before it (.cpp:495-503)
    g.second_mate_.resize(R);
    g.read_weights_.resize(R);
    for (int r = 0; r < R; r++) {
      g.second_mate_[r] = (r > 0 && rd->name_id[r0 + r] == rd->name_id[r0 + r - 1]) ? 1 : 0;
      g.read_weights_[r] = g.second_mate_[r] ? 0 : 1;
    }
    g.read_cigar_off_.resize(R + 1);
    for (int r = 0; r <= R; r++) g.read_cigar_off_[r] = rd->cigar_off[r0 + r] - rd->cigar_off[r0];
    g.read_cigar_type_.assign(rd->cigar_type + rd->cigar_off[r0], rd->cigar_type + rd->cigar_off[r1]);
    g.read_cigar_len_.assign(rd->cigar_len + rd->cigar_off[r0], rd->cigar_len + rd->cigar_off[r1]);
    // ReadPooler: identical sequences share a pool, upper-median qualities
    std::vector<int32_t> seq_off(R + 1);
    for (int r = 0; r <= R; r++) seq_off[r] = rd->read_seq_off[r0 + r] - rd->read_seq_off[r0];
    const char* bases = rd->bases + rd->read_seq_off[r0];
    const char* quals = rd->quals + rd->read_seq_off[r0];
    g.read_seq_off_ = seq_off;
    g.read_quals_.assign(quals, quals + seq_off[R]);
    g.pool_index_.resize(R);
    std::vector<int32_t> first(std::max(R, 1));
    g.pool_seq_off_.resize(R + 1);
    g.pool_bases_.resize(seq_off[R]);
    g.pool_quals_.resize(seq_off[R]);
    int32_t P = 0;
    hipstr_status_t st = hipstr_pool_reads(R, seq_off.data(), bases, quals, g.pool_index_.data(), &P, first.data(),
                                           g.pool_seq_off_.data(), &g.pool_bases_[0], &g.pool_quals_[0]);
    if (st != HIPSTR_OK) { err = "hipstr_pool_reads failed"; return st; }
    g.num_pools_ = P;
    g.pool_seq_off_.resize(P + 1);
    g.pool_bases_.resize(g.pool_seq_off_[P]);
    g.pool_quals_.resize(g.pool_seq_off_[P]);
    // seeds from the pool's first member alignment (HapAligner::calc_seed_base)
    std::vector<int32_t> starts(P), lens(P), coff(P + 1, 0), clen, rs, re;
    std::vector<char> ctype;
    for (int p = 0; p < P; p++) {
      const int r = first[p];
      starts[p] = g.read_start_[r];
      lens[p] = seq_off[r + 1] - seq_off[r];
      ctype.insert(ctype.end(), g.read_cigar_type_.begin() + g.read_cigar_off_[r], g.read_cigar_type_.begin() + g.read_cigar_off_[r + 1]);
      clen.insert(clen.end(), g.read_cigar_len_.begin() + g.read_cigar_off_[r], g.read_cigar_len_.begin() + g.read_cigar_off_[r + 1]);
      coff[p + 1] = (int32_t)ctype.size();
    }
    for (const HapBlock& b : g.hap_blocks_)
      if (b.period > 0) { rs.push_back(b.start); re.push_back(b.end); }
    g.pool_seed_.assign(P, -1);
    if (!g.hap_blocks_.empty())
      st = hipstr_calc_seeds(P, starts.data(), lens.data(), coff.data(), ctype.data(), clen.data(), g.hap_blocks_.front().start,
                             g.hap_blocks_.back().end, (int32_t)rs.size(), rs.data(), re.data(), g.pool_seed_.data());
    if (st != HIPSTR_OK) { err = "hipstr_calc_seeds failed (the reference dies on a seed at a read end)"; return st; }
    g.seed_positions_.assign(R, -1);
    g.sample_total_LLs_.assign(g.num_samples_, 0.0);
    g.optimal_haps_.assign((size_t)g.num_samples_ * 2, 0);
    if (!g.hap_blocks_.empty()) g.rebuild_hap_aln_info(nullptr);
  }
  return HIPSTR_OK;
}

hipstr_status_t GenotyperBatch::add_loci(const hipstr_align_batch_t* bt, const int32_t* block_start, const int32_t* block_end,
                                         const hipstr_locus_reads_t* rd, std::string& err) {
  if (!bt || !block_start || !block_end || !rd) { err = "null argument"; return HIPSTR_ERR_BAD_ARG; }
  for (int l = 0; l < bt->n_loci; l++) {
    loci.emplace_back();
    SeqStutterGenotyper& g = loci.back();
    const int b0 = bt->locus_block_off[l], b1 = bt->locus_block_off[l + 1];
    if (b1 - b0 < 1 || b1 - b0 > HIPSTR_MAX_BLOCKS_PER_LOCUS) { err = "unsupported number of haplotype blocks"; return HIPSTR_ERR_BAD_ARG; }
    g.num_alleles_ = 1;
    for (int b = b0; b < b1; b++) {
      HapBlock blk;
      blk.start = block_start[b]; blk.end = block_end[b]; blk.period = bt->block_period[b];
      std::memcpy(blk.stutter, bt->block_stutter + 6 * (size_t)b, sizeof(blk.stutter));
      for (int o = bt->block_opt_off[b]; o < bt->block_opt_off[b + 1]; o++)
        blk.seqs.emplace_back(bt->opt_seq + bt->opt_seq_off[o], bt->opt_seq + bt->opt_seq_off[o + 1]);
      if (blk.seqs.empty()) { err = "haplotype block without a reference allele"; return HIPSTR_ERR_BAD_ARG; }
      g.num_alleles_ *= blk.num_options();
      g.hap_blocks_.push_back(blk);
    }
    hipstr_status_t st = init_reads(g, rd, l, err);
    if (st != HIPSTR_OK) return st;
  }
  return HIPSTR_OK;
}

hipstr_status_t GenotyperBatch::add_loci_from_reads(int32_t n_loci, const int32_t* region_start, const int32_t* region_stop,
                                                    const int32_t* period, const char* const* chrom_seq, const double* stutter,
                                                    const hipstr_locus_reads_t* rd, std::string& err, const int32_t* allele_pos,
                                                    const int32_t* allele_off, const char* const* alleles) {
  if (!region_start || !region_stop || !period || !chrom_seq || !stutter || !rd) { err = "null argument"; return HIPSTR_ERR_BAD_ARG; }
  const size_t base = loci.size();
  loci.resize(base + n_loci);
  std::vector<hipstr_status_t> status(n_loci, HIPSTR_OK);
  std::vector<std::string> errors(n_loci);
  host_tables();
  // Loci of one chromosome share its sequence: the length is measured once per distinct pointer and the sequence is used
  // in place (a copy per locus would move the whole chromosome -- hundreds of MB -- for every locus).
  std::vector<size_t> chrom_len(n_loci, 0);
  {
    std::unordered_map<const char*, size_t> seen;
    for (int l = 0; l < n_loci; l++) {
      if (!chrom_seq[l]) { err = "null chromosome sequence"; return HIPSTR_ERR_BAD_ARG; }
      auto it = seen.find(chrom_seq[l]);
      if (it == seen.end()) it = seen.emplace(chrom_seq[l], std::strlen(chrom_seq[l])).first;
      chrom_len[l] = it->second;
    }
  }
  parallel_for(n_loci, [&](size_t li) {
    const int l = (int)li;
    SeqStutterGenotyper& g = loci[base + li];
    // build_haplotype (seq_stutter_genotyper.cpp:422-484): blocks from the reads that span the padded region
    const int r0 = rd->locus_read_off[l], r1 = rd->locus_read_off[l + 1];
    const int S = rd->locus_sample_off[l + 1] - rd->locus_sample_off[l];
    std::vector<std::vector<ReadView> > by_sample(S);
    int32_t min_start = INT32_MAX, max_stop = INT32_MIN;
    for (int r = r0; r < r1; r++) {
      ReadView v;
      v.start = rd->read_start[r];
      v.bases = rd->bases + rd->read_seq_off[r];
      v.n_cigar = rd->cigar_off[r + 1] - rd->cigar_off[r];
      v.cigar_type = rd->cigar_type + rd->cigar_off[r];
      v.cigar_len = rd->cigar_len + rd->cigar_off[r];
      if (rd->read_stop) v.stop = rd->read_stop[r];
      else {   // HipSTR's convention: the last aligned reference position
        v.stop = v.start - 1;
        for (int c = 0; c < v.n_cigar; c++)
          if (v.cigar_type[c] != 'I') v.stop += v.cigar_len[c];
      }
      min_start = std::min(min_start, v.start);
      max_stop = std::max(max_stop, v.stop);
      if (!rd->use_for_haps || rd->use_for_haps[r]) by_sample[rd->sample_label[r]].push_back(v);
    }
    g.log_ += "Generating candidate haplotypes\n";
    const std::string_view chrom(chrom_seq[l], chrom_len[li]);
    HaplotypeGenerator generator(min_start, max_stop);
    bool added;
    if (allele_pos) {   // ref_vcf != NULL (seq_stutter_genotyper.cpp:445-459): the alleles of the reference panel's record
      g.fixed_alleles_ = true;
      std::vector<std::string> vcf_alleles;
      for (int a = allele_off[l]; a < allele_off[l + 1]; a++) vcf_alleles.push_back(alleles[a]);
      if (allele_pos[l] < 0 || vcf_alleles.empty()) {
        g.log_ += "Haplotype construction failed: The alleles could not be extracted from the reference VCF\n";
        g.phase_ = SeqStutterGenotyper::FAILED;
        status[li] = init_reads(g, rd, l, errors[li]);
        return;
      }
      added = generator.add_vcf_haplotype_block(allele_pos[l], period[l], chrom, vcf_alleles, stutter + 6 * (size_t)l);
    } else added = generator.add_haplotype_block(region_start[l], region_stop[l], period[l], chrom, by_sample, stutter + 6 * (size_t)l);
    if (added && generator.fuse_haplotype_blocks(chrom)) {
      g.hap_blocks_ = generator.get_haplotype_blocks();
      g.num_alleles_ = 1;
      for (const HapBlock& b : g.hap_blocks_) g.num_alleles_ *= b.num_options();
    } else {
      g.log_ += "Haplotype construction failed: " + generator.failure_msg() + "\n";
      g.phase_ = SeqStutterGenotyper::FAILED;   // initialized_ = false: genotype() returns false
    }
    status[li] = init_reads(g, rd, l, errors[li]);
  });
  for (int l = 0; l < n_loci; l++)
    if (status[l] != HIPSTR_OK) { err = errors[l]; return status[l]; }
  return HIPSTR_OK;
}

namespace {

/* Uninitialised buffer from the pluggable host allocator (flatten.h): page-locked through the C-ABI layer's block cache
 * when a context exists, so the copies to and from the device run at PCIe speed; std::vector would also zero (and
 * page-fault) hundreds of MB on one core before the workers fill it. */
template <class T>
struct RawBuf {
  T* p = nullptr;
  size_t n = 0;
  RawBuf() {}
  RawBuf(const RawBuf&) = delete;
  RawBuf& operator=(const RawBuf&) = delete;
  ~RawBuf() { if (p) host_free(p); }
  void alloc(size_t count) { if (p) host_free(p); n = count; p = static_cast<T*>(host_alloc(std::max<size_t>(count, 1) * sizeof(T))); }
  T* data() { return p; }
};

}  // namespace

/* One masked K1 + K2 + K3 call for the loci in `which`.  Packing a window of loci into the flat batch touches ~0.7 GB for
 * 1 000 loci; it is done by all host threads at once (offsets by prefix sums first, then every locus copies its slices). */
hipstr_status_t GenotyperBatch::run_alignments(const std::vector<int>& which, std::string& err) {
  if (which.empty()) return HIPSTR_OK;
  const double t_pack = now_s();
  const size_t L = which.size();
  std::vector<SeqStutterGenotyper*> gs(L);
  for (size_t k = 0; k < L; k++) gs[k] = &loci[which[k]];
  // prefix sums
  std::vector<int32_t> locus_block_off(L + 1, 0), locus_pool_off(L + 1, 0), locus_read_off(L + 1, 0), locus_sample_off(L + 1, 0),
      locus_opt_off(L + 1, 0), n_haps(L);
  std::vector<int64_t> locus_hap_off(L + 1, 0), locus_out_off(L + 1, 0), opt_byte_off(L + 1, 0), pool_byte_off(L + 1, 0), ll_off(L + 1, 0),
      post_off(L + 1, 0);
  std::vector<uint8_t> haploid(L);
  bool pool_masked = false, hap_masked = false, copy_masked = false;
  for (size_t k = 0; k < L; k++) {
    const SeqStutterGenotyper& g = *gs[k];
    int opts = 0;
    int64_t opt_bytes = 0;
    for (const HapBlock& b : g.hap_blocks_) {
      opts += b.num_options();
      for (const std::string& q : b.seqs) opt_bytes += (int64_t)q.size();
    }
    locus_block_off[k + 1] = locus_block_off[k] + (int32_t)g.hap_blocks_.size();
    locus_opt_off[k + 1] = locus_opt_off[k] + opts;
    opt_byte_off[k + 1] = opt_byte_off[k] + opt_bytes;
    locus_pool_off[k + 1] = locus_pool_off[k] + g.num_pools_;
    pool_byte_off[k + 1] = pool_byte_off[k] + (int64_t)g.pool_bases_.size();
    locus_hap_off[k + 1] = locus_hap_off[k] + g.num_alleles_;
    locus_out_off[k + 1] = locus_out_off[k] + (int64_t)g.num_pools_ * g.num_alleles_;
    locus_read_off[k + 1] = locus_read_off[k] + g.num_reads_;
    locus_sample_off[k + 1] = locus_sample_off[k] + g.num_samples_;
    ll_off[k + 1] = ll_off[k] + (int64_t)g.num_reads_ * g.num_alleles_;
    post_off[k + 1] = post_off[k] + (int64_t)g.num_samples_ * g.num_alleles_ * g.num_alleles_;
    n_haps[k] = g.num_alleles_;
    haploid[k] = g.haploid_ ? 1 : 0;
    for (uint8_t m : g.realign_hap_) hap_masked |= (m == 0);
    for (uint8_t m : g.realign_pool_) pool_masked |= (m == 0);
    copy_masked |= !g.copy_read_.empty();
  }
  if (pool_byte_off[L] > INT32_MAX) { err = "window too large: pooled read bytes exceed the 32-bit offsets of the batch"; return HIPSTR_ERR_BAD_ARG; }
  const bool masked = pool_masked || hap_masked || copy_masked;   // only then must the previous likelihoods travel
  const size_t B = locus_block_off[L], O = locus_opt_off[L], P = locus_pool_off[L], R = locus_read_off[L], S = locus_sample_off[L];
  RawBuf<int32_t> block_period, block_opt_off, opt_seq_off, pool_seq_off, pool_seed, pool_index, sample_label, read_weight, read_seed, best;
  RawBuf<double> block_stutter, log_p1, log_p2, read_ll, post, sample_ll, total_ll;
  RawBuf<char> opt_seq, pool_bases, pool_quals;
  RawBuf<uint8_t> realign_pool, realign_hap, second_mate, copy_read;
  block_period.alloc(B); block_opt_off.alloc(B + 1); block_stutter.alloc(6 * B); opt_seq_off.alloc(O + 1); opt_seq.alloc(opt_byte_off[L]);
  pool_seq_off.alloc(P + 1); pool_seed.alloc(P); pool_bases.alloc(pool_byte_off[L]); pool_quals.alloc(pool_byte_off[L]);
  realign_pool.alloc(P); realign_hap.alloc(locus_hap_off[L]);
  pool_index.alloc(R); sample_label.alloc(R); read_weight.alloc(R); second_mate.alloc(R); copy_read.alloc(R); log_p1.alloc(R); log_p2.alloc(R);
  read_ll.alloc(ll_off[L]); read_seed.alloc(R); post.alloc(post_off[L]); sample_ll.alloc(S); best.alloc(2 * S); total_ll.alloc(L);
  block_opt_off.p[0] = 0; opt_seq_off.p[0] = 0; pool_seq_off.p[0] = 0;
  parallel_for(L, [&](size_t k) {
    const SeqStutterGenotyper& g = *gs[k];
    int32_t b = locus_block_off[k], o = locus_opt_off[k];
    int64_t ob = opt_byte_off[k];
    for (const HapBlock& blk : g.hap_blocks_) {
      block_period.p[b] = blk.period;
      std::memcpy(block_stutter.p + 6 * (size_t)b, blk.stutter, 6 * sizeof(double));
      for (const std::string& q : blk.seqs) {
        std::memcpy(opt_seq.p + ob, q.data(), q.size());
        ob += (int64_t)q.size();
        opt_seq_off.p[++o] = (int32_t)ob;
      }
      block_opt_off.p[++b] = o;
    }
    const int32_t p0 = locus_pool_off[k];
    const int64_t pb0 = pool_byte_off[k];
    std::memcpy(pool_bases.p + pb0, g.pool_bases_.data(), g.pool_bases_.size());
    std::memcpy(pool_quals.p + pb0, g.pool_quals_.data(), g.pool_quals_.size());
    for (int q = 0; q < g.num_pools_; q++) {
      pool_seq_off.p[p0 + q + 1] = (int32_t)(pb0 + g.pool_seq_off_[q + 1]);
      pool_seed.p[p0 + q] = g.pool_seed_[q];
      realign_pool.p[p0 + q] = g.realign_pool_.empty() ? 1 : g.realign_pool_[q];
    }
    for (int h = 0; h < g.num_alleles_; h++) realign_hap.p[locus_hap_off[k] + h] = g.realign_hap_.empty() ? 1 : g.realign_hap_[h];
    const int32_t r0 = locus_read_off[k];
    const size_t nr = (size_t)g.num_reads_;
    std::memcpy(pool_index.p + r0, g.pool_index_.data(), nr * sizeof(int32_t));
    std::memcpy(sample_label.p + r0, g.sample_label_.data(), nr * sizeof(int32_t));
    std::memcpy(read_weight.p + r0, g.read_weights_.data(), nr * sizeof(int32_t));
    std::memcpy(second_mate.p + r0, g.second_mate_.data(), nr);
    std::memcpy(log_p1.p + r0, g.log_p1_.data(), nr * sizeof(double));
    std::memcpy(log_p2.p + r0, g.log_p2_.data(), nr * sizeof(double));
    if (g.copy_read_.empty()) std::memset(copy_read.p + r0, 1, nr);
    else std::memcpy(copy_read.p + r0, g.copy_read_.data(), nr);
    if (masked) {   // in-place semantics of log_aln_probs_ / seed_positions_ under the masks
      std::memcpy(read_ll.p + ll_off[k], g.log_aln_probs_.data(), (size_t)(ll_off[k + 1] - ll_off[k]) * sizeof(double));
      std::memcpy(read_seed.p + r0, g.seed_positions_.data(), nr * sizeof(int32_t));
    } else {        // results only: fault the pages in here, on all threads, not one by one under the device-to-host copy
      for (int64_t i = ll_off[k]; i < ll_off[k + 1]; i += 512) read_ll.p[i] = 0.0;
      for (size_t i = 0; i < nr; i += 1024) read_seed.p[r0 + i] = 0;
    }
    for (int64_t i = post_off[k]; i < post_off[k + 1]; i += 512) post.p[i] = 0.0;
  });
  hipstr_align_batch_t bt;
  std::memset(&bt, 0, sizeof(bt));
  bt.n_loci = (int32_t)L; bt.n_blocks = (int32_t)B; bt.n_options = (int32_t)O; bt.n_pools = (int32_t)P; bt.n_haps = locus_hap_off[L];
  bt.locus_block_off = locus_block_off.data(); bt.locus_pool_off = locus_pool_off.data();
  bt.locus_hap_off = locus_hap_off.data(); bt.locus_out_off = locus_out_off.data();
  bt.block_period = block_period.p; bt.block_opt_off = block_opt_off.p; bt.block_stutter = block_stutter.p;
  bt.opt_seq_off = opt_seq_off.p; bt.opt_seq = opt_seq.p;
  bt.pool_seq_off = pool_seq_off.p; bt.pool_bases = pool_bases.p; bt.pool_quals = pool_quals.p; bt.pool_seed = pool_seed.p;
  bt.realign_pool = pool_masked ? realign_pool.p : nullptr;
  bt.realign_hap = hap_masked ? realign_hap.p : nullptr;
  hipstr_reads_batch_t rb;
  rb.locus_read_off = locus_read_off.data(); rb.locus_sample_off = locus_sample_off.data();
  rb.pool_index = pool_index.p; rb.sample_label = sample_label.p; rb.second_mate = second_mate.p; rb.read_weight = read_weight.p;
  rb.log_p1 = log_p1.p; rb.log_p2 = log_p2.p; rb.haploid = haploid.data();
  rb.copy_read = copy_masked ? copy_read.p : nullptr;
  hipstr_genotype_out_t out;
  out.read_ll = read_ll.p; out.read_seed = read_seed.p; out.post = post.p; out.sample_ll = sample_ll.p; out.best = best.p;
  out.total_ll = total_ll.p;
  seconds[T_ALIGN_PACK] += now_s() - t_pack;
  hipstr_status_t st = hipstr_genotype_batch_host(ctx_, &bt, &rb, &out);
  account_device_call();
  if (st != HIPSTR_OK) { err = std::string("hipstr_genotype_batch_host: ") + hipstr_last_error(ctx_); return st; }
  n_alignments += hipstr_batch_num_alignments(&bt);
  const double t_unpack = now_s();
  parallel_for(L, [&](size_t k) {
    SeqStutterGenotyper& g = *gs[k];
    g.log_aln_probs_.assign(read_ll.p + ll_off[k], read_ll.p + ll_off[k + 1]);
    g.seed_positions_.assign(read_seed.p + locus_read_off[k], read_seed.p + locus_read_off[k + 1]);
    g.log_sample_posteriors_.assign(post.p + post_off[k], post.p + post_off[k + 1]);
    g.sample_total_LLs_.assign(sample_ll.p + locus_sample_off[k], sample_ll.p + locus_sample_off[k + 1]);
    g.optimal_haps_.assign(best.p + 2 * (size_t)locus_sample_off[k], best.p + 2 * (size_t)locus_sample_off[k + 1]);
  });
  seconds[T_ALIGN_UNPACK] += now_s() - t_unpack;
  return HIPSTR_OK;
}

/* One K3 call for the loci in `which` (loci that only lost alleles: the likelihoods are unchanged, only the posteriors are
 * recomputed).  Packed by all host threads into page-locked buffers like run_alignments. */
hipstr_status_t GenotyperBatch::run_posteriors(const std::vector<int>& which, std::string& err) {
  if (which.empty()) return HIPSTR_OK;
  const size_t L = which.size();
  std::vector<SeqStutterGenotyper*> gs(L);
  for (size_t k = 0; k < L; k++) gs[k] = &loci[which[k]];
  std::vector<int32_t> locus_read_off(L + 1, 0), locus_sample_off(L + 1, 0), n_haps(L);
  std::vector<int64_t> ll_off(L + 1, 0), post_off(L + 1, 0);
  std::vector<uint8_t> haploid(L);
  for (size_t k = 0; k < L; k++) {
    const SeqStutterGenotyper& g = *gs[k];
    locus_read_off[k + 1] = locus_read_off[k] + g.num_reads_;
    locus_sample_off[k + 1] = locus_sample_off[k] + g.num_samples_;
    ll_off[k + 1] = ll_off[k] + (int64_t)g.num_reads_ * g.num_alleles_;
    post_off[k + 1] = post_off[k] + (int64_t)g.num_samples_ * g.num_alleles_ * g.num_alleles_;
    n_haps[k] = g.num_alleles_;
    haploid[k] = g.haploid_ ? 1 : 0;
  }
  const size_t R = locus_read_off[L], S = locus_sample_off[L];
  RawBuf<int32_t> sample_label, read_weight, best;
  RawBuf<double> log_p1, log_p2, read_ll, post, sample_ll, total_ll;
  sample_label.alloc(R); read_weight.alloc(R); log_p1.alloc(R); log_p2.alloc(R); read_ll.alloc(ll_off[L]);
  post.alloc(post_off[L]); sample_ll.alloc(S); best.alloc(2 * S); total_ll.alloc(L);
  parallel_for(L, [&](size_t k) {
    const SeqStutterGenotyper& g = *gs[k];
    const int32_t r0 = locus_read_off[k];
    const size_t nr = (size_t)g.num_reads_;
    std::memcpy(sample_label.p + r0, g.sample_label_.data(), nr * sizeof(int32_t));
    std::memcpy(read_weight.p + r0, g.read_weights_.data(), nr * sizeof(int32_t));
    std::memcpy(log_p1.p + r0, g.log_p1_.data(), nr * sizeof(double));
    std::memcpy(log_p2.p + r0, g.log_p2_.data(), nr * sizeof(double));
    std::memcpy(read_ll.p + ll_off[k], g.log_aln_probs_.data(), (size_t)(ll_off[k + 1] - ll_off[k]) * sizeof(double));
  });
  hipstr_status_t st = hipstr_posteriors_host(ctx_, (int32_t)L, locus_read_off.data(), locus_sample_off.data(), n_haps.data(),
                                              haploid.data(), read_ll.p, log_p1.p, log_p2.p, sample_label.p, read_weight.p, post.p,
                                              sample_ll.p, best.p, total_ll.p);
  account_device_call();
  if (st != HIPSTR_OK) { err = std::string("hipstr_posteriors_host: ") + hipstr_last_error(ctx_); return st; }
  parallel_for(L, [&](size_t k) {
    SeqStutterGenotyper& g = *gs[k];
    g.log_sample_posteriors_.assign(post.p + post_off[k], post.p + post_off[k + 1]);
    g.sample_total_LLs_.assign(sample_ll.p + locus_sample_off[k], sample_ll.p + locus_sample_off[k + 1]);
    g.optimal_haps_.assign(best.p + 2 * (size_t)locus_sample_off[k], best.p + 2 * (size_t)locus_sample_off[k + 1]);
  });
  return HIPSTR_OK;
}

hipstr_status_t GenotyperBatch::run_traces(const std::vector<int>& which, std::string& err) {
  const size_t kChunk = 1 << 17;   // traces per device call (bounds the host result buffers)
  size_t li = 0, ti = 0;           // next locus of `which`, next missing trace of that locus
  // result buffers of one chunk, allocated once for the widest stride of the call and reused by every chunk (fresh
  // memory would be page-faulted in again, serially, ~0.8 KB per trace)
  size_t total = 0;
  int32_t widest_read = 0, widest_hap = 0;   // over ALL loci of the call: a chunk mixes loci, and the stride must hold its longest read plus its longest haplotype
  for (int l : which) {
    const SeqStutterGenotyper& g = loci[l];
    if (g.missing_traces_.empty()) continue;
    total += g.missing_traces_.size();
    int32_t rd = 0, hp = 0;
    for (int p = 0; p < g.num_pools_; p++) rd = std::max(rd, g.pool_seq_off_[p + 1] - g.pool_seq_off_[p]);
    for (const HapBlock& b : g.hap_blocks_) {
      size_t m = 0;
      for (const auto& q : b.seqs) m = std::max(m, q.size());
      hp += (int32_t)m;
    }
    widest_read = std::max(widest_read, rd);
    widest_hap = std::max(widest_hap, hp);
  }
  if (total == 0) { for (int l : which) { loci[l].missing_traces_.clear(); loci[l].missing_trace_read_.clear(); } return HIPSTR_OK; }
  const int32_t stride = ((widest_read + widest_hap + 2 + 15) / 16) * 16;
  const size_t cap = std::min(total, kChunk);
  RawBuf<char> hap_aln;
  RawBuf<int32_t> seed_hap_pos, stutter, span_start, span_len, flank_ins, flank_del, n_indels, indels, n_snps, snps;
  hap_aln.alloc(cap * (size_t)stride);
  seed_hap_pos.alloc(cap); stutter.alloc(cap * 8); span_start.alloc(cap * 8); span_len.alloc(cap * 8); flank_ins.alloc(cap);
  flank_del.alloc(cap); n_indels.alloc(cap); indels.alloc(cap * HIPSTR_MAX_TRACE_INDELS * 2); n_snps.alloc(cap);
  snps.alloc(cap * HIPSTR_MAX_TRACE_SNPS * 2);
  while (li < which.size()) {
    PackedBatch pb;
    std::vector<int32_t> trace_pool, trace_hap;
    std::vector<std::pair<int, int> > owner;   // (locus, index into missing_traces_)
    while (li < which.size() && trace_pool.size() < kChunk) {
      SeqStutterGenotyper& g = loci[which[li]];
      if (ti >= g.missing_traces_.size()) { li++; ti = 0; continue; }
      if (ti == 0) g.trace_cache_.reserve(g.trace_cache_.size() + g.missing_traces_.size());
      const int32_t pool_base = (int32_t)pb.pool_seed.size();
      std::vector<int> own;   // reads traced with their own qualities become extra pools of this locus
      if (!g.missing_trace_read_.empty())
        for (size_t t = ti; t < g.missing_traces_.size() && trace_pool.size() + (t - ti) < kChunk; t++) own.push_back(g.missing_trace_read_[t]);
      pb.add(g, nullptr, nullptr, &own);
      for (size_t first = ti; ti < g.missing_traces_.size() && trace_pool.size() < kChunk; ti++) {
        trace_pool.push_back(pool_base + (own.empty() ? g.missing_traces_[ti].first : g.num_pools_ + (int)(ti - first)));
        trace_hap.push_back(g.missing_traces_[ti].second);
        owner.emplace_back(which[li], (int)ti);
      }
      if (ti >= g.missing_traces_.size()) { li++; ti = 0; }
    }
    const size_t n = trace_pool.size();
    if (n == 0) break;
    hipstr_trace_out_t out;
    out.aln_stride = stride;
    out.hap_aln = hap_aln.data();
    out.seed_hap_pos = seed_hap_pos.data();
    out.stutter_size = stutter.data();
    out.span_start = span_start.data();
    out.span_len = span_len.data();
    out.flank_ins = flank_ins.data();
    out.flank_del = flank_del.data();
    out.n_indels = n_indels.data();
    out.indels = indels.data();
    out.n_snps = n_snps.data();
    out.snps = snps.data();
    hipstr_align_batch_t bt = pb.view();
    const double t_dev = now_s();
    hipstr_status_t st = hipstr_trace_batch_host(ctx_, &bt, pb.block_start.data(), (int32_t)n, trace_pool.data(), trace_hap.data(), &out);
    account_device_call();
    seconds[T_TRACE_DEVICE] += now_s() - t_dev;
    if (st != HIPSTR_OK) { err = std::string("hipstr_trace_batch_host: ") + hipstr_last_error(ctx_); return st; }
    n_traces += (int64_t)n;
    // stitch every trace against the reference and file it in its locus' cache; traces of one locus are contiguous
    std::vector<size_t> group_start;
    for (size_t i = 0; i < n; i++)
      if (i == 0 || owner[i].first != owner[i - 1].first) group_start.push_back(i);
    group_start.push_back(n);
    std::atomic<int> failed(0);
    parallel_for(group_start.size() - 1, [&](size_t grp) {
      std::vector<char> ctype(stride + 8), aln(2 * (size_t)stride + 8);
      std::vector<int32_t> clen(stride + 8);
      for (size_t i = group_start[grp]; i < group_start[grp + 1]; i++) {
      SeqStutterGenotyper& g = loci[owner[i].first];
      const std::pair<int, int> key = g.missing_traces_[owner[i].second];
      const int nb = (int)g.hap_blocks_.size();
      const std::string_view read = g.pool_read_view(key.first);
      AlignmentTrace& t = g.trace_cache_[key];
      t = AlignmentTrace();
      const char* ops = hap_aln.p + i * (size_t)stride;   // read-vs-haplotype operations, NUL-terminated
      t.flank_ins_size = flank_ins.p[i];
      t.flank_del_size = flank_del.p[i];
      t.num_blocks = nb;
      for (int b = 0; b < nb; b++) {
        t.stutter_size[b] = stutter.p[i * 8 + b];
        t.block_seq[b] = span_len.p[i * 8 + b] > 0 ? read.substr(span_start.p[i * 8 + b], span_len.p[i * 8 + b]) : std::string_view();
      }
      if (n_indels.p[i] > HIPSTR_MAX_TRACE_INDELS || n_snps.p[i] > HIPSTR_MAX_TRACE_SNPS) {
        // more flank indels / SNPs than the fixed slots of the device call hold (a chimeric or mismapped read): the counts
        // are the true ones, the complete lists are rebuilt on the host from the trace's operation string
        std::vector<int32_t> all_indels(2 * (size_t)std::max(n_indels.p[i], 1)), all_snps(2 * (size_t)std::max(n_snps.p[i], 1));
        int32_t ni = 0, ns = 0;
        const hipstr_status_t st3 = hipstr_trace_flank_lists(&bt, pb.block_start.data(), trace_pool[i], trace_hap[i], ops,
                                                             seed_hap_pos.p[i], stutter.p + i * 8, nullptr, n_indels.p[i], &ni,
                                                             all_indels.data(), n_snps.p[i], &ns, all_snps.data());
        if (st3 != HIPSTR_OK || ni != n_indels.p[i] || ns != n_snps.p[i]) { failed = 1; return; }
        for (int k = 0; k < ni; k++) t.flank_indel_data.emplace_back(all_indels[2 * (size_t)k], all_indels[2 * (size_t)k + 1]);
        for (int k = 0; k < ns; k++) t.flank_snp_data.emplace_back(all_snps[2 * (size_t)k], (char)all_snps[2 * (size_t)k + 1]);
      } else {
        for (int k = 0; k < n_indels.p[i]; k++)
          t.flank_indel_data.emplace_back(indels.p[(i * HIPSTR_MAX_TRACE_INDELS + k) * 2], indels.p[(i * HIPSTR_MAX_TRACE_INDELS + k) * 2 + 1]);
        for (int k = 0; k < n_snps.p[i]; k++)
          t.flank_snp_data.emplace_back(snps.p[(i * HIPSTR_MAX_TRACE_SNPS + k) * 2], (char)snps.p[(i * HIPSTR_MAX_TRACE_SNPS + k) * 2 + 1]);
      }
      // only the span against the reference is needed by the loop and the VCF record; the CIGAR / gapped string of the
      // traced alignment (used by the reference's HTML visualisation) are built on request (keep_traced_alignments)
      int32_t n_cigar = 0;
      const bool full = keep_traced_alignments;
      const std::string& to_ref = g.hap_aln_info_[key.second];
      const std::vector<int32_t>& to_ref_index = g.hap_aln_index_[key.second];
      const hipstr_status_t st2 =
          (!full && !to_ref_index.empty())
              ? hipstr_trace_span(g.hap_blocks_.front().start, to_ref.c_str(), (int32_t)to_ref.size(), to_ref_index.data(), ops,
                                  seed_hap_pos.p[i], g.pool_seed_[key.first], &t.start, &t.stop)
              : hipstr_stitch_trace(g.hap_blocks_.front().start, to_ref.c_str(), ops, seed_hap_pos.p[i], g.pool_seed_[key.first],
                                    full ? std::string(read).c_str() : "", &t.start, &t.stop, (int32_t)ctype.size(),
                                    full ? ctype.data() : nullptr, full ? clen.data() : nullptr, &n_cigar, (int32_t)aln.size(),
                                    full ? aln.data() : nullptr);
      if (st2 != HIPSTR_OK) { failed = 1; return; }
      if (full) {
        std::ostringstream cig;
        for (int k = 0; k < n_cigar; k++) cig << clen[k] << ctype[k];
        t.cigar = cig.str();
        t.alignment = std::string(aln.data());
        t.hap_aln = ops;
      }
      }
    });
    if (failed) { err = "hipstr_stitch_trace failed"; return HIPSTR_ERR_BAD_ARG; }
  }
  for (int l : which) { loci[l].missing_traces_.clear(); loci[l].missing_trace_read_.clear(); }
  return HIPSTR_OK;
}

hipstr_status_t GenotyperBatch::genotype(int max_total_haplotypes, int max_flank_haplotypes, double min_flank_freq,
                                         bool reassemble_flanks, std::string& err) {
  if (!ctx_) { err = "no device context: the genotyping loop only runs on the GPU"; return HIPSTR_ERR_NO_DEVICE; }
  const int kMinKmer = 10, kMaxKmer = 15;   // seq_stutter_genotyper.h:153-154
  for (SeqStutterGenotyper& g : loci) {
    if (g.phase_ != SeqStutterGenotyper::ALIGN_ALL) continue;
    g.max_total_haplotypes_ = max_total_haplotypes;
    g.max_flank_haplotypes_ = max_flank_haplotypes;
    g.min_flank_freq_ = min_flank_freq;
    g.reassemble_flanks_ = reassemble_flanks;
    if (g.num_alleles_ > max_total_haplotypes) {
      std::ostringstream msg;
      msg << "Aborting genotyping of the locus as too many candidate haplotypes were found (# Found = " << g.num_alleles_
          << ", MAX = " << max_total_haplotypes << ")\n";
      g.log_ += msg.str();
      g.phase_ = SeqStutterGenotyper::FAILED;
      continue;
    }
    // flanks too repetitive to assemble -> skip the locus (seq_stutter_genotyper.cpp:616-630)
    for (int flank = 0; flank < 2 && g.phase_ != SeqStutterGenotyper::FAILED; flank++) {
      const std::string& ref_seq = (flank == 0 ? g.hap_blocks_.front() : g.hap_blocks_.back()).seqs[0];
      const int max_k = std::min(kMaxKmer, ref_seq.empty() ? -1 : (int)ref_seq.size() - 1);
      int k = 0;
      if (!FlankAssembler::calc_kmer_length(ref_seq, kMinKmer, max_k, k)) {
        g.log_ += std::string("Aborting genotyping of the locus as the sequence ") + (flank == 0 ? "upstream" : "downstream") +
                  " of the repeat is too repetitive for accurate genotyping\n";
        g.phase_ = SeqStutterGenotyper::FAILED;
      }
    }
  }
  for (;;) {
    std::vector<int> need_traces, need_alignment, need_posteriors;
    const double t_decide = now_s();
    std::vector<SeqStutterGenotyper::Request> requests(loci.size());
    host_tables();   // build the constant tables before the workers read them
    parallel_for(loci.size(), [&](size_t l) { requests[l] = loci[l].advance(); });
    for (size_t l = 0; l < loci.size(); l++) {
      switch (requests[l]) {
        case SeqStutterGenotyper::NEED_TRACES: need_traces.push_back((int)l); break;
        case SeqStutterGenotyper::NEED_ALIGNMENT: need_alignment.push_back((int)l); break;
        case SeqStutterGenotyper::NEED_POSTERIORS: need_posteriors.push_back((int)l); break;
        case SeqStutterGenotyper::NONE: break;
      }
    }
    seconds[T_DECIDE] += now_s() - t_decide;
    if (need_traces.empty() && need_alignment.empty() && need_posteriors.empty()) break;
    n_rounds++;
    double t0 = now_s();
    const double dev0 = seconds[T_TRACE_DEVICE];
    hipstr_status_t st = run_traces(need_traces, err);
    double t1 = now_s();
    seconds[T_TRACE_HOST] += (t1 - t0) - (seconds[T_TRACE_DEVICE] - dev0);
    if (st == HIPSTR_OK) st = run_alignments(need_alignment, err);
    t0 = now_s();
    seconds[T_ALIGN] += t0 - t1;
    if (st == HIPSTR_OK) st = run_posteriors(need_posteriors, err);
    seconds[T_POSTERIORS] += now_s() - t0;
    if (st != HIPSTR_OK) return st;
  }
  return HIPSTR_OK;
}

hipstr_status_t GenotyperBatch::recompute_stutter_models(int max_total_haplotypes, int max_flank_haplotypes, double min_flank_freq,
                                                         int max_em_iter, double abs_ll_converge, double frac_ll_converge,
                                                         std::string& err) {
  if (!ctx_) { err = "no device context"; return HIPSTR_ERR_NO_DEVICE; }
  std::vector<int> which;
  for (size_t l = 0; l < loci.size(); l++)
    if (loci[l].succeeded()) which.push_back((int)l);
  if (which.empty()) return HIPSTR_OK;
  // retrace_alignments for the final genotypes
  std::vector<int> need;
  for (int l : which) {
    loci[l].log_ += "Retraining EM stutter genotyper using maximum likelihood alignments\n";
    if (!loci[l].collect_missing_traces()) need.push_back(l);
  }
  hipstr_status_t st = run_traces(need, err);
  if (st != HIPSTR_OK) return st;
  // one EM problem per (locus, repeat block)
  std::vector<int32_t> read_off{0}, sample_off{0}, num_bps, labels, motif, ref_allele;
  std::vector<double> p1, p2;
  std::vector<uint8_t> haploid;
  std::vector<std::pair<int, int> > problem;   // (locus, block)
  for (int l : which) {
    SeqStutterGenotyper& g = loci[l];
    for (int b = 0; b < (int)g.hap_blocks_.size(); b++) {
      const HapBlock& block = g.hap_blocks_[b];
      if (block.period <= 0) continue;
      for (int r = 0; r < g.num_reads_; r++) {
        if (g.seed_positions_[r] < 0) continue;
        const AlignmentTrace& t = g.trace_cache_.at(std::make_pair(g.pool_index_[r], g.best_hap_of_read(r)));
        if (!(t.start < block.start && t.stop > block.end)) continue;
        num_bps.push_back((int32_t)t.str_seq(b).size() + t.stutter_size[b]);
        labels.push_back(g.sample_label_[r]);
        p1.push_back(g.log_p1_[r]);
        p2.push_back(g.log_p2_[r]);
      }
      read_off.push_back((int32_t)num_bps.size());
      sample_off.push_back(sample_off.back() + g.num_samples_);
      motif.push_back(block.period);
      ref_allele.push_back(0);   // the reference passes 0 as the reference allele size (.cpp:1570)
      haploid.push_back(g.haploid_ ? 1 : 0);
      problem.emplace_back(l, b);
    }
  }
  hipstr_em_batch_t em;
  em.n_loci = (int32_t)problem.size();
  em.locus_read_off = read_off.data();
  em.locus_sample_off = sample_off.data();
  em.num_bps = num_bps.data();
  em.sample_label = labels.data();
  em.log_p1 = p1.data();
  em.log_p2 = p2.data();
  em.motif_len = motif.data();
  em.ref_allele = ref_allele.data();
  em.haploid = haploid.data();
  std::vector<double> params(6 * problem.size()), ll(problem.size());
  std::vector<uint8_t> converged(problem.size());
  std::vector<int32_t> iters(problem.size());
  st = hipstr_em_train_host(ctx_, &em, max_em_iter, abs_ll_converge, frac_ll_converge, params.data(), converged.data(), iters.data(),
                            ll.data());
  account_device_call();
  if (st != HIPSTR_OK) { err = std::string("hipstr_em_train_host: ") + hipstr_last_error(ctx_); return st; }
  for (size_t k = 0; k < problem.size(); k++) {
    SeqStutterGenotyper& g = loci[problem[k].first];
    if (g.phase_ == SeqStutterGenotyper::FAILED) continue;
    if (!converged[k]) {
      g.log_ += "Retraining stutter model training failed\n";
      g.phase_ = SeqStutterGenotyper::FAILED;
      continue;
    }
    std::memcpy(g.hap_blocks_[problem[k].second].stutter, &params[6 * k], 6 * sizeof(double));
  }
  for (int l : which) {
    SeqStutterGenotyper& g = loci[l];
    if (g.phase_ == SeqStutterGenotyper::FAILED) continue;
    g.trace_cache_.clear();
    g.phase_ = SeqStutterGenotyper::ALIGN_ALL;   // genotype() again, from the current allele set
  }
  return genotype(max_total_haplotypes, max_flank_haplotypes, min_flank_freq, loci[which[0]].reassemble_flanks_, err);
}

}  // namespace hipstr

/* ---- C-ABI ------------------------------------------------------------------------------------------ */
extern "C" {

hipstr_status_t hipstr_hap_aln_to_ref(const char* ref_hap, const char* alt_hap, int32_t first_block_start,
                                      int32_t repeat_block_start, int32_t cap, char* out) {
  if (!ref_hap || !alt_hap || !out) return HIPSTR_ERR_BAD_ARG;
  const std::string info = hipstr::hap_aln_to_ref(ref_hap, alt_hap, first_block_start, repeat_block_start);
  if (info.empty() || (int32_t)info.size() + 1 > cap) return HIPSTR_ERR_BAD_ARG;
  std::memcpy(out, info.c_str(), info.size() + 1);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_create(hipstr_ctx_t* ctx, const hipstr_align_batch_t* blocks, const int32_t* block_start,
                                        const int32_t* block_end, const hipstr_locus_reads_t* reads, hipstr_genotyper_t** out) {
  if (!out) return HIPSTR_ERR_BAD_ARG;   // ctx may be NULL: construction is host work, genotype() then needs a device
  hipstr_genotyper* g = new hipstr_genotyper(ctx);
  const double t0 = hipstr::now_s();
  hipstr_status_t st = g->batch.add_loci(blocks, block_start, block_end, reads, g->last_error);
  g->batch.seconds[hipstr::GenotyperBatch::T_CONSTRUCT] += hipstr::now_s() - t0;
  if (st != HIPSTR_OK) { delete g; return st; }
  *out = g;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_create_from_reads(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* region_start,
                                                   const int32_t* region_stop, const int32_t* period, const char* const* chrom_seq,
                                                   const double* stutter, const hipstr_locus_reads_t* reads, hipstr_genotyper_t** out) {
  if (!out || n_loci < 0) return HIPSTR_ERR_BAD_ARG;
  hipstr_genotyper* g = new hipstr_genotyper(ctx);
  const double t0 = hipstr::now_s();
  hipstr_status_t st = g->batch.add_loci_from_reads(n_loci, region_start, region_stop, period, chrom_seq, stutter, reads, g->last_error);
  g->batch.seconds[hipstr::GenotyperBatch::T_CONSTRUCT] += hipstr::now_s() - t0;
  if (st != HIPSTR_OK) { delete g; return st; }
  *out = g;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_create_with_ref_alleles(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* region_start,
                                                         const int32_t* region_stop, const int32_t* period, const char* const* chrom_seq,
                                                         const double* stutter, const hipstr_locus_reads_t* reads, const int32_t* allele_pos,
                                                         const int32_t* allele_off, const char* const* alleles, hipstr_genotyper_t** out) {
  if (!out || n_loci < 0 || !allele_pos || !allele_off || (allele_off[n_loci] > 0 && !alleles)) return HIPSTR_ERR_BAD_ARG;
  hipstr_genotyper* g = new hipstr_genotyper(ctx);
  const double t0 = hipstr::now_s();
  hipstr_status_t st = g->batch.add_loci_from_reads(n_loci, region_start, region_stop, period, chrom_seq, stutter, reads, g->last_error, allele_pos,
                                                    allele_off, alleles);
  g->batch.seconds[hipstr::GenotyperBatch::T_CONSTRUCT] += hipstr::now_s() - t0;
  if (st != HIPSTR_OK) { delete g; return st; }
  *out = g;
  return HIPSTR_OK;
}

void hipstr_genotyper_destroy(hipstr_genotyper_t* g) { delete g; }
const char* hipstr_genotyper_last_error(const hipstr_genotyper_t* g) { return g ? g->last_error.c_str() : "null genotyper"; }

hipstr_status_t hipstr_genotyper_genotype(hipstr_genotyper_t* g, int32_t max_total_haplotypes, int32_t max_flank_haplotypes,
                                          double min_flank_freq, int32_t reassemble_flanks, uint8_t* locus_ok) {
  if (!g) return HIPSTR_ERR_BAD_ARG;
  hipstr_status_t st = g->batch.genotype(max_total_haplotypes, max_flank_haplotypes, min_flank_freq, reassemble_flanks != 0,
                                         g->last_error);
  if (st != HIPSTR_OK) return st;
  if (locus_ok)
    for (size_t l = 0; l < g->batch.loci.size(); l++) locus_ok[l] = g->batch.loci[l].succeeded() ? 1 : 0;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_recompute_stutter_models(hipstr_genotyper_t* g, int32_t max_total_haplotypes,
                                                          int32_t max_flank_haplotypes, double min_flank_freq, int32_t max_em_iter,
                                                          double abs_ll_converge, double frac_ll_converge, uint8_t* locus_ok) {
  if (!g) return HIPSTR_ERR_BAD_ARG;
  hipstr_status_t st = g->batch.recompute_stutter_models(max_total_haplotypes, max_flank_haplotypes, min_flank_freq, max_em_iter,
                                                         abs_ll_converge, frac_ll_converge, g->last_error);
  if (st != HIPSTR_OK) return st;
  if (locus_ok)
    for (size_t l = 0; l < g->batch.loci.size(); l++) locus_ok[l] = g->batch.loci[l].succeeded() ? 1 : 0;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_timing(const hipstr_genotyper_t* g, double* seconds9) {
  if (!g || !seconds9) return HIPSTR_ERR_BAD_ARG;
  for (int i = 0; i < hipstr::GenotyperBatch::T_COUNT; i++) seconds9[i] = g->batch.seconds[i];
  return HIPSTR_OK;
}
/* host seconds of the per-locus decisions summed over loci, by phase: {align-all set-up, stutter-allele discovery,
 * uncalled pruning, unspanned pruning, flank assembly, post-assembly pruning, done, failed} */
hipstr_status_t hipstr_genotyper_phase_timing(const hipstr_genotyper_t* g, double* seconds8) {
  if (!g || !seconds8) return HIPSTR_ERR_BAD_ARG;
  for (int i = 0; i < 8; i++) seconds8[i] = 0;
  for (const auto& l : g->batch.loci)
    for (int i = 0; i < 8; i++) seconds8[i] += l.phase_seconds_[i];
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_stats(const hipstr_genotyper_t* g, int64_t* n_alignments, int64_t* n_traces, int32_t* n_rounds) {
  if (!g) return HIPSTR_ERR_BAD_ARG;
  if (n_alignments) *n_alignments = g->batch.n_alignments;
  if (n_traces) *n_traces = g->batch.n_traces;
  if (n_rounds) *n_rounds = g->batch.n_rounds;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_locus_info(const hipstr_genotyper_t* g, int32_t locus, int32_t* info) {
  if (!g || !info || locus < 0 || locus >= (int32_t)g->batch.loci.size()) return HIPSTR_ERR_BAD_ARG;
  const hipstr::SeqStutterGenotyper& s = g->batch.loci[locus];
  int32_t n_opts = 0, seq_bytes = 0;
  for (const auto& b : s.hap_blocks_) {
    n_opts += b.num_options();
    for (const auto& q : b.seqs) seq_bytes += (int32_t)q.size();
  }
  info[0] = (int32_t)s.hap_blocks_.size();
  info[1] = s.num_alleles_;
  info[2] = s.num_reads_;
  info[3] = s.num_samples_;
  info[4] = s.num_pools_;
  info[5] = n_opts;
  info[6] = seq_bytes;
  info[7] = s.rounds_;
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_locus_blocks(const hipstr_genotyper_t* g, int32_t locus, int32_t* block_n_opts,
                                              int32_t* opt_seq_off, char* opt_seq) {
  if (!g || !block_n_opts || !opt_seq_off || !opt_seq || locus < 0 || locus >= (int32_t)g->batch.loci.size()) return HIPSTR_ERR_BAD_ARG;
  const hipstr::SeqStutterGenotyper& s = g->batch.loci[locus];
  int32_t o = 0, at = 0;
  opt_seq_off[0] = 0;
  for (size_t b = 0; b < s.hap_blocks_.size(); b++) {
    block_n_opts[b] = s.hap_blocks_[b].num_options();
    for (const auto& q : s.hap_blocks_[b].seqs) {
      std::memcpy(opt_seq + at, q.data(), q.size());
      at += (int32_t)q.size();
      opt_seq_off[++o] = at;
    }
  }
  return HIPSTR_OK;
}

hipstr_status_t hipstr_genotyper_locus_results(const hipstr_genotyper_t* g, int32_t locus, double* read_ll, int32_t* read_seed,
                                               int32_t* pool_index, double* post, double* sample_ll, int32_t* best,
                                               uint8_t* call_sample_ok) {
  if (!g || locus < 0 || locus >= (int32_t)g->batch.loci.size()) return HIPSTR_ERR_BAD_ARG;
  const hipstr::SeqStutterGenotyper& s = g->batch.loci[locus];
  if (read_ll) std::copy(s.log_aln_probs_.begin(), s.log_aln_probs_.end(), read_ll);
  if (read_seed) std::copy(s.seed_positions_.begin(), s.seed_positions_.end(), read_seed);
  if (pool_index) std::copy(s.pool_index_.begin(), s.pool_index_.end(), pool_index);
  if (post) std::copy(s.log_sample_posteriors_.begin(), s.log_sample_posteriors_.end(), post);
  if (sample_ll) std::copy(s.sample_total_LLs_.begin(), s.sample_total_LLs_.end(), sample_ll);
  if (best) std::copy(s.optimal_haps_.begin(), s.optimal_haps_.end(), best);
  if (call_sample_ok)
    for (int i = 0; i < s.num_samples_; i++) call_sample_ok[i] = s.call_sample_[i].empty() ? 1 : 0;
  return HIPSTR_OK;
}

int32_t hipstr_genotyper_locus_log(const hipstr_genotyper_t* g, int32_t locus, char* out, int32_t cap) {
  if (!g || !out || cap <= 0 || locus < 0 || locus >= (int32_t)g->batch.loci.size()) return -1;
  const std::string& s = g->batch.loci[locus].log_;
  const size_t n = std::min(s.size(), (size_t)cap - 1);
  std::memcpy(out, s.data() + (s.size() - n), n);
  out[n] = 0;
  return (int32_t)n;
}

}  // extern "C"
