/*
 * vcf_record.cpp -- a18: SeqStutterGenotyper::write_vcf_record (src/seq_stutter_genotyper.cpp:984-1510) for a batch of
 * loci.  The numeric inputs come from the device (K3b: genotype posteriors, GLs, PLs; K5: traces of the reads against
 * their strand-assigned haplotype); this file holds the reference's bookkeeping and text formatting.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

#include "../csrc/flatten.h"
#include "seq_stutter_genotyper.h"
#include "vcf_writer.h"

namespace hipstr {

const double kTolerance = 1e-10;        // mathops.cpp:10
const double kStrandTolerance = 0.1;    // seq_stutter_genotyper.h:157

std::string fixed2(double v) {          // out.precision(2); out.setf(std::ios::fixed)
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.2f", v);
  return buf;
}
std::string upper(std::string s) {
  for (char& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}
double log_sum_exp2(double a, double b) {   // mathops.cpp:52-57
  return a > b ? a + std::log(1 + std::exp(b - a)) : b + std::log(1 + std::exp(a - b));
}
double log_sum_exp(const std::vector<double>& v) {   // mathops.cpp:64-70
  const double m = *std::max_element(v.begin(), v.end());
  double total = 0;
  for (double x : v) total += std::exp(x - m);
  return m + std::log(total);
}

/* key|count pairs of a list of bp differences, e.g. "-4|2;0|11" (Genotyper::condense_read_counts) */
std::string condense_read_counts(const std::vector<int>& diffs) {
  if (diffs.empty()) return ".";
  std::map<int, int> counts;
  for (int d : diffs) counts[d]++;
  std::ostringstream out;
  for (auto it = counts.begin(); it != counts.end(); ++it) {
    if (it != counts.begin()) out << ";";
    out << it->first << "|" << it->second;
  }
  return out.str();
}

/* Base-pair difference of a read from the reference inside [region_start, region_end], read off its CIGAR
 * (ExtractCigar, extract_indels.cpp:18-90): the window is widened to the nearest match operations and the read
 * must cover it entirely. */
bool extract_cigar(const char* type, const int32_t* len, int n, int cigar_start, int region_start, int region_end, int& bp_diff) {
  auto is_match = [](char t) { return t == 'M' || t == '=' || t == 'X'; };
  int ref_span = 0;
  for (int i = 0; i < n; i++)
    if (is_match(type[i]) || type[i] == 'D') ref_span += len[i];
  if (region_start < cigar_start || region_end >= cigar_start + ref_span) return false;
  int pos = cigar_start, first = 0, last_match = 0;
  while (pos < region_start && first < n) {
    if (is_match(type[first]) || type[first] == 'D') pos += len[first];
    if (is_match(type[first])) last_match = first;
    first++;
  }
  first = last_match;
  if (first == 0 && !is_match(type[0])) return false;
  int last = n - 1;
  last_match = n - 1;
  pos = cigar_start + ref_span;
  while (pos > region_end) {
    if (is_match(type[last]) || type[last] == 'D') pos -= len[last];
    if (is_match(type[last])) last_match = last;
    if (last == 0) break;
    last--;
  }
  last = last_match;
  if (last == n - 1 && !is_match(type[last])) return false;
  bp_diff = 0;
  for (int i = first; i <= last; i++)
    if (type[i] == 'D') bp_diff -= len[i];
    else if (type[i] == 'I') bp_diff += len[i];
  return true;
}

/* lgamma(n + 1) for read counts: the very values std::lgamma returns, computed once (three calls per binomial coefficient and a
 * few dozen coefficients per sample were 2 % of the loop's host time) */
double lgamma_of_count_plus_one(int n) {
  static const std::vector<double> table = [] {
    std::vector<double> t(4096);
    for (int i = 0; i < (int)t.size(); i++) t[i] = std::lgamma(i + 1.0);
    return t;
  }();
  return (n >= 0 && n < (int)table.size()) ? table[n] : std::lgamma(n + 1.0);
}
double log_binomial(int n, int k) {
  return (k == 0 || n == k) ? 0.0 : lgamma_of_count_plus_one(n) - lgamma_of_count_plus_one(k) - lgamma_of_count_plus_one(n - k);
}

/* log10 of the two-sided binomial p-value for the read split between the two haplotypes
 * (compute_allele_bias, seq_stutter_genotyper.cpp:965-982; the reference takes the CDF from cephes' bdtr). */
double compute_allele_bias(int hap_a_reads, int hap_b_reads) {
  const int total = hap_a_reads + hap_b_reads;
  if (total == 0) return 1;
  if (hap_a_reads == hap_b_reads) return 0.0;
  const int k = std::min(hap_a_reads, hap_b_reads);
  double cdf = 0;
  for (int j = 0; j <= k; j++) cdf += std::exp(log_binomial(total, j) + total * std::log(0.5));
  return std::log10(std::min(1.0, 2 * cdf));
}

/* Two-sided p-value of Fisher's exact test on a 2x2 table, as htslib 1.9's kt_fisher_exact defines it
 * (lib/htslib/kfunc.c:196-279): the sum over both tails of the tables at most as probable as the observed one,
 * with its 1e-8 relative slack. */
double fisher_two_sided(int n11, int n12, int n21, int n22) {
  const int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
  const int hi = std::min(n_1, n1_), lo = std::max(0, n1_ + n_1 - n);
  if (lo == hi) return 1.0;
  auto pmf = [&](int x) { return std::exp(log_binomial(n1_, x) + log_binomial(n - n1_, n_1 - x) - log_binomial(n, n_1)); };
  const double q = pmf(n11);
  double left = 0, right = 0, p = pmf(lo);
  int i = lo + 1;
  for (; p < 0.99999999 * q && i <= hi; ++i) { left += p; p = pmf(i); }
  if (p < 1.00000001 * q) left += p;
  p = pmf(hi);
  int j = hi - 1;
  for (; p < 0.99999999 * q && j >= 0; --j) { right += p; p = pmf(j); }
  if (p < 1.00000001 * q) right += p;
  return std::min(1.0, left + right);
}

/* Alleles of the STR block as the VCF reports them: trimmed to the region where all alleles agree, padded back with
 * reference sequence, one base added on the left when an allele would be empty or start differently
 * (get_alleles, seq_stutter_genotyper.cpp:691-769). */
std::pair<int, int> get_alleles(const HapBlock& block, int32_t region_start, int32_t region_stop, std::string_view chrom_seq,
                                int32_t& pos, std::vector<std::string>& alleles) {
  alleles = block.seqs;
  int32_t left_trim = 0, start = block.start;
  while (start + left_trim < region_start) {
    bool trim = true;
    for (const std::string& a : alleles)
      if ((size_t)(left_trim + 1) >= a.size() || a[left_trim] != alleles[0][left_trim]) { trim = false; break; }
    if (!trim) break;
    left_trim++;
  }
  start += left_trim;
  for (std::string& a : alleles) a = a.substr(left_trim);
  int32_t right_trim = 0, end = block.end;
  while (end - right_trim > region_stop) {
    bool trim = true;
    const int ref_size = (int)alleles[0].size();
    for (const std::string& a : alleles) {
      const int alt_size = (int)a.size();
      if ((size_t)(right_trim + 1) >= a.size() || a[alt_size - right_trim - 1] != alleles[0][ref_size - right_trim - 1]) { trim = false; break; }
    }
    if (!trim) break;
    right_trim++;
  }
  end -= right_trim;
  for (std::string& a : alleles) a = a.substr(0, a.size() - right_trim);
  std::string left_flank = start >= region_start ? upper(std::string(chrom_seq.substr(region_start, start - region_start))) : "";
  const std::string right_flank = end <= region_stop ? upper(std::string(chrom_seq.substr(end, region_stop - end))) : "";
  pos = std::min(region_start, start);
  left_trim -= (int32_t)left_flank.size();
  right_trim -= (int32_t)right_flank.size();
  if (left_flank.empty()) {
    bool pad_left = false;
    for (size_t i = 1; i < alleles.size(); i++)
      if (alleles[i].empty() || alleles[i][0] != alleles[0][0]) { pad_left = true; break; }
    if (pad_left) {
      pos -= 1;
      left_trim -= 1;
      left_flank = upper(std::string(chrom_seq.substr(pos, 1)));
    }
  }
  for (std::string& a : alleles) a = left_flank + a + right_flank;
  pos += 1;   // VCF positions are 1-based
  return std::make_pair(left_trim, right_trim);
}

/* reference allele first, the others by (length, sequence) (reorder_alleles, seq_stutter_genotyper.cpp:673-689) */
void reorder_alleles(const std::vector<std::string>& alleles, std::vector<int>& old_to_new, std::vector<int>& new_to_old) {
  std::map<std::string, int> old_index;
  for (size_t i = 0; i < alleles.size(); i++) old_index[alleles[i]] = (int)i;
  std::vector<std::string> sorted = alleles;
  std::sort(sorted.begin() + 1, sorted.end(), [](const std::string& a, const std::string& b) {
    return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
  });
  old_to_new.assign(alleles.size(), -1);
  new_to_old.clear();
  for (size_t i = 0; i < sorted.size(); i++) {
    const int o = old_index[sorted[i]];
    new_to_old.push_back(o);
    old_to_new[o] = (int)i;
  }
}

/* K3b outputs of one locus (views into the batch-wide arrays) */
struct LocusGenotypes {
  const int32_t *best_hap, *best_gt, *pl;
  const double *log_phased, *log_unphased, *hap_log_phased, *hap_log_unphased, *gl, *phased_gl, *gl_diff;
  int n_gl, n_phased_gl;   // entries per sample
};

void SeqStutterGenotyper::vcf_prepare(const int32_t* best_hap) {
  const double log_one_half = host_tables().log_one_half;
  vcf_calls_.assign(num_reads_, ReadCall{0, 0, 0.0, false});
  missing_traces_.clear();
  missing_trace_read_.clear();
  std::set<std::pair<int, int> > wanted;
  for (int r = 0; r < num_reads_; r++) {
    if (seed_positions_[r] < 0) continue;
    const int s = sample_label_[r], hap_a = best_hap[2 * s], hap_b = best_hap[2 * s + 1];
    const double* ll = &log_aln_probs_[(size_t)r * num_alleles_];
    const double total = log_sum_exp2(log_one_half + log_p1_[r] + ll[hap_a], log_one_half + log_p2_[r] + ll[hap_b]);
    ReadCall& c = vcf_calls_[r];
    c.log_phase_one = log_one_half + log_p1_[r] + ll[hap_a] - total;
    if (!haploid_ && (hap_a != hap_b || std::fabs(log_p1_[r] - log_p2_[r]) > kTolerance)) {
      const double v1 = log_p1_[r] + ll[hap_a], v2 = log_p2_[r] + ll[hap_b];
      if (std::fabs(v1 - v2) > kStrandTolerance) { c.read_strand = v1 > v2 ? 0 : 1; c.unique = true; }
    }
    c.best_hap = c.read_strand == 0 ? hap_a : hap_b;
    const std::pair<int, int> key(pool_index_[r], c.best_hap);
    // the first read that needs an uncached trace is traced with its own qualities (.cpp:1116-1120)
    if (trace_cache_.count(key) == 0 && wanted.insert(key).second) { missing_traces_.push_back(key); missing_trace_read_.push_back(r); }
  }
}

namespace {

void format_record(SeqStutterGenotyper& g, const LocusGenotypes& k3b, const std::string& chrom, const std::string& name,
                   int32_t region_start, int32_t region_stop, int32_t period, std::string_view chrom_seq,
                   const std::vector<std::string>& locus_names, const std::vector<std::string>& out_names,
                   const hipstr_vcf_options_t& opt) {
  const int S = g.num_samples_, nb = (int)g.hap_blocks_.size();
  int hap_block_index = -1;
  for (int b = 0; b < nb; b++)
    if (g.hap_blocks_[b].period > 0) { hap_block_index = b; break; }
  const HapBlock& block = g.hap_blocks_[hap_block_index];
  int32_t pos;
  std::vector<std::string> alleles;
  const std::pair<int, int> trimmings = get_alleles(block, region_start, region_stop, chrom_seq, pos, alleles);

  // flank alleles adjusted by what get_alleles moved between the flanks and the repeat
  std::vector<std::string> lflank_seqs, rflank_seqs;
  std::vector<int> hap_to_lflank, hap_to_rflank;
  if (opt.output_haplotype_data && nb == 3) {
    const std::string& ref_str = g.hap_blocks_[1].seqs[0];
    const int ref_len = (int)ref_str.size();
    g.haps_to_alleles(0, hap_to_lflank);
    for (const std::string& seq : g.hap_blocks_[0].seqs)
      lflank_seqs.push_back(trimmings.first < 0 ? seq.substr(0, (int)seq.size() + trimmings.first) : seq + ref_str.substr(0, trimmings.first));
    g.haps_to_alleles(2, hap_to_rflank);
    for (const std::string& seq : g.hap_blocks_[2].seqs)
      rflank_seqs.push_back(trimmings.second < 0 ? seq.substr(trimmings.second)
                                                 : ref_str.substr(ref_len - trimmings.second, trimmings.second) + seq);
  }
  std::vector<int> allele_bp_diffs;
  for (const std::string& a : alleles) allele_bp_diffs.push_back((int)a.size() - (int)alleles[0].size());
  std::vector<int> hap_to_allele;
  g.haps_to_alleles(hap_block_index, hap_to_allele);

  // per-read bookkeeping, grouped by sample
  std::vector<int> num_aligned(S, 0), num_snps(S, 0), num_stutter(S, 0), num_flank_indels(S, 0), strand_one(S, 0), strand_two(S, 0),
      unique_one(S, 0), unique_two(S, 0), rv_unique_one(S, 0), rv_unique_two(S, 0);
  std::vector<std::vector<int> > bps(S), ml_bps(S);
  std::vector<std::vector<double> > log_read_phases(S);
  for (int r = 0; r < g.num_reads_; r++) {
    if (g.seed_positions_[r] < 0) continue;
    const int s = g.sample_label_[r];
    const SeqStutterGenotyper::ReadCall& c = g.vcf_calls_[r];
    log_read_phases[s].push_back(c.log_phase_one);
    if (c.unique) {
      (c.read_strand == 0 ? unique_one : unique_two)[s]++;
      if (g.rev_strand_[r]) (c.read_strand == 0 ? rv_unique_one : rv_unique_two)[s]++;
    }
    const AlignmentTrace& trace = g.trace_cache_.at(std::make_pair(g.pool_index_[r], c.best_hap));
    if (trace.has_stutter()) num_stutter[s]++;
    if (trace.flank_ins_size != 0 || trace.flank_del_size != 0) num_flank_indels[s]++;
    num_aligned[s]++;
    if (std::fabs(g.log_p1_[r] - g.log_p2_[r]) > kTolerance) {
      num_snps[s]++;
      (g.log_p1_[r] > g.log_p2_[r] ? strand_one : strand_two)[s]++;
    }
    int bp_diff;
    const int c0 = g.read_cigar_off_[r], c1 = g.read_cigar_off_[r + 1];
    if (extract_cigar(&g.read_cigar_type_[c0], &g.read_cigar_len_[c0], c1 - c0, g.read_start_[r], region_start - period,
                      region_stop + period, bp_diff))
      bps[s].push_back(bp_diff);
    if (trace.start < (region_start > 4 ? region_start - 4 : 0) && trace.stop > region_stop + 4)
      ml_bps[s].push_back(allele_bp_diffs[hap_to_allele[c.best_hap]] + trace.total_stutter_size());
  }

  std::map<std::string, int> sample_indices;
  for (int s = 0; s < S; s++) sample_indices.insert(std::make_pair(locus_names[s], s));
  const std::set<std::string> samples_of_interest(out_names.begin(), out_names.end());
  auto too_many_flank_indels = [&](int s) { return num_aligned[s] > 0 && num_flank_indels[s] > (float)opt.max_flank_indel_frac * num_aligned[s]; };

  std::vector<int> allele_counts(alleles.size(), 0);
  int skip_count = 0, filt_count = 0, allele_number = 0;
  for (int s = 0; s < S; s++) {
    if (!samples_of_interest.count(locus_names[s]) || num_aligned[s] == 0) continue;
    if (too_many_flank_indels(s)) { filt_count++; continue; }
    if (!g.call_sample_[s].empty()) { skip_count++; continue; }
    allele_counts[k3b.best_gt[2 * s]]++;
    allele_number++;
    if (!g.haploid_) { allele_counts[k3b.best_gt[2 * s + 1]]++; allele_number++; }
  }
  std::vector<int> old_to_new, new_to_old;
  reorder_alleles(alleles, old_to_new, new_to_old);

  std::ostringstream out;
  out << chrom << "\t" << pos << "\t" << (name.empty() ? "." : name) << "\t" << alleles[new_to_old[0]] << "\t";
  if (alleles.size() == 1) out << ".";
  else {
    for (size_t i = 1; i + 1 < alleles.size(); i++) out << alleles[new_to_old[i]] << ",";
    out << alleles[new_to_old.back()];
  }
  out << "\t.\t.";
  // StutterModel::get_parameter order: in-frame geom / up / down, out-of-frame geom / up / down
  out << "\tINFRAME_PGEOM=" << fixed2(block.stutter[0]) << ";INFRAME_UP=" << fixed2(block.stutter[1]) << ";INFRAME_DOWN="
      << fixed2(block.stutter[2]) << ";OUTFRAME_PGEOM=" << fixed2(block.stutter[3]) << ";OUTFRAME_UP=" << fixed2(block.stutter[4])
      << ";OUTFRAME_DOWN=" << fixed2(block.stutter[5]) << ";START=" << region_start + 1 << ";END=" << region_stop << ";PERIOD="
      << period << ";NSKIP=" << skip_count << ";NFILT=" << filt_count << ";";
  if (alleles.size() > 1) {
    out << "BPDIFFS=" << allele_bp_diffs[new_to_old[1]];
    for (size_t i = 2; i < alleles.size(); i++) out << "," << allele_bp_diffs[new_to_old[i]];
    out << ";";
  }
  int tot_dp = 0, tot_dsnp = 0, tot_dstutter = 0, tot_dflankindel = 0;
  for (const std::string& nm : out_names) {
    auto it = sample_indices.find(nm);
    if (it == sample_indices.end()) continue;
    const int s = it->second;
    if (!g.call_sample_[s].empty() || too_many_flank_indels(s)) continue;
    tot_dp += num_aligned[s];
    tot_dsnp += num_snps[s];
    tot_dstutter += num_stutter[s];
    tot_dflankindel += num_flank_indels[s];
  }
  out << "DP=" << tot_dp << ";DSNP=" << tot_dsnp << ";DSTUTTER=" << tot_dstutter << ";DFLANKINDEL=" << tot_dflankindel << ";";
  out << "AN=" << allele_number << ";REFAC=" << allele_counts[0];
  if (allele_counts.size() > 1) {
    out << ";AC=";
    for (size_t i = 1; i + 1 < allele_counts.size(); i++) out << allele_counts[new_to_old[i]] << ",";
    out << allele_counts[new_to_old.back()];
  }
  bool output_lflanks = false, output_rflanks = false;
  if (opt.output_haplotype_data) {
    if (lflank_seqs.size() > 1) {
      output_lflanks = true;
      out << ";LFLANKS=" << lflank_seqs[0];
      for (size_t i = 1; i < lflank_seqs.size(); i++) out << "," << lflank_seqs[i];
    }
    if (rflank_seqs.size() > 1) {
      output_rflanks = true;
      out << ";RFLANKS=" << rflank_seqs[0];
      for (size_t i = 1; i < rflank_seqs.size(); i++) out << "," << rflank_seqs[i];
    }
  }
  const bool output_bias = !g.haploid_ && g.reassemble_flanks();   // AB / DAB / FS need all reads and the assembly
  int num_fields;
  if (!g.haploid_) { out << "\tGT:GB:Q:PQ:DP:DSNP:DSTUTTER:DFLANKINDEL:PDP:PSNP:GLDIFF"; num_fields = 11; }
  else { out << "\tGT:GB:Q:DP:DSTUTTER:DFLANKINDEL:GLDIFF"; num_fields = 7; }
  if (output_bias) out << ":AB:DAB:FS";
  if (opt.output_allreads) out << ":ALLREADS";
  if (opt.output_mallreads) out << ":MALLREADS";
  if (opt.output_gls) out << ":GL";
  if (opt.output_pls) out << ":PL";
  if (!g.haploid_ && opt.output_phased_gls) out << ":PHASEDGL";
  if (opt.output_haplotype_data)
    out << (output_lflanks || output_rflanks ? ":HQ:PHQ" : "") << (output_lflanks ? ":LFGT" : "") << (output_rflanks ? ":RFGT" : "");
  if (opt.output_filters) out << ":FILTER";
  num_fields += (output_bias ? 3 : 0) + (!g.haploid_ && opt.output_phased_gls ? 1 : 0);
  num_fields += (opt.output_allreads ? 1 : 0) + (opt.output_mallreads ? 1 : 0) + (opt.output_gls ? 1 : 0) + (opt.output_pls ? 1 : 0) +
                (output_lflanks || output_rflanks ? 2 : 0) + (output_lflanks ? 1 : 0) + (output_rflanks ? 1 : 0);
  std::string empty_str;
  for (int n = 0; n < num_fields; n++) empty_str += ".:";
  auto missing = [&](const std::string& why) { return opt.output_filters ? empty_str + why : std::string("."); };

  std::map<std::string, int> filter_reasons;
  const int V = (int)new_to_old.size();
  for (const std::string& nm : out_names) {
    out << "\t";
    auto it = sample_indices.find(nm);
    if (it == sample_indices.end()) { out << missing("NO_READS"); continue; }
    const int s = it->second;
    if (num_aligned[s] == 0) { filter_reasons["NO_READS"]++; out << missing("NO_READS"); continue; }
    if (!g.call_sample_[s].empty()) { filter_reasons[g.call_sample_[s]]++; out << missing(g.call_sample_[s]); continue; }
    if (too_many_flank_indels(s)) {
      g.call_sample_[s] = "FLANK_INDEL_FRAC";
      filter_reasons["FLANK_INDEL_FRAC"]++;
      out << missing("FLANK_INDEL_FRAC");
      continue;
    }
    const double phase1_reads = std::exp(log_sum_exp(log_read_phases[s])), phase2_reads = num_aligned[s] - phase1_reads;
    const int gt_a = k3b.best_gt[2 * s], gt_b = k3b.best_gt[2 * s + 1], hap_a = k3b.best_hap[2 * s], hap_b = k3b.best_hap[2 * s + 1];
    double allele_bias = 1.01, strand_bias = 1.01;
    if (!g.haploid_ && hap_a != hap_b) {
      allele_bias = compute_allele_bias(unique_one[s], unique_two[s]);
      strand_bias = std::log10(std::min(1.0, fisher_two_sided(unique_one[s] - rv_unique_one[s], rv_unique_one[s],
                                                               unique_two[s] - rv_unique_two[s], rv_unique_two[s])));
    }
    const double* gls = k3b.gl + (size_t)s * k3b.n_gl;
    const int32_t* pls = k3b.pl + (size_t)s * k3b.n_gl;
    const double* phased_gls = k3b.phased_gl + (size_t)s * k3b.n_phased_gl;
    if (!g.haploid_) {
      out << old_to_new[gt_a] << "|" << old_to_new[gt_b] << ":" << allele_bp_diffs[gt_a] << "|" << allele_bp_diffs[gt_b] << ":"
          << fixed2(std::exp(k3b.log_unphased[s])) << ":" << fixed2(std::exp(k3b.log_phased[s])) << ":" << num_aligned[s] << ":"
          << num_snps[s] << ":" << num_stutter[s] << ":" << num_flank_indels[s] << ":" << fixed2(phase1_reads) << "|"
          << fixed2(phase2_reads) << ":" << strand_one[s] << "|" << strand_two[s];
    } else {
      out << old_to_new[gt_a] << ":" << allele_bp_diffs[gt_a] << ":" << fixed2(std::exp(k3b.log_unphased[s])) << ":" << num_aligned[s]
          << ":" << num_stutter[s] << ":" << num_flank_indels[s];
    }
    out << ":" << (alleles.size() == 1 ? std::string(".") : fixed2(k3b.gl_diff[s]));
    if (output_bias) {
      if (allele_bias > 1) out << ":0:.";
      else out << ":" << fixed2(allele_bias) << ":" << unique_one[s] + unique_two[s];
      if (strand_bias > 1) out << ":0";
      else out << ":" << fixed2(strand_bias);
    }
    if (opt.output_allreads) out << ":" << condense_read_counts(bps[s]);
    if (opt.output_mallreads) out << ":" << condense_read_counts(ml_bps[s]);
    // likelihood fields follow the new allele order
    auto diploid_index = [&](int i, int j) {
      const int a = std::min(new_to_old[i], new_to_old[j]), b = std::max(new_to_old[i], new_to_old[j]);
      return b * (b + 1) / 2 + a;
    };
    if (g.haploid_) {
      if (opt.output_gls) { out << ":" << fixed2(gls[0]); for (int i = 1; i < V; i++) out << "," << fixed2(gls[new_to_old[i]]); }
      if (opt.output_pls) { out << ":" << pls[0]; for (int i = 1; i < V; i++) out << "," << pls[new_to_old[i]]; }
    } else {
      if (opt.output_gls) {
        out << ":" << fixed2(gls[0]);
        for (int i = 1; i < V; i++) for (int j = 0; j <= i; j++) out << "," << fixed2(gls[diploid_index(i, j)]);
      }
      if (opt.output_pls) {
        out << ":" << pls[0];
        for (int i = 1; i < V; i++) for (int j = 0; j <= i; j++) out << "," << pls[diploid_index(i, j)];
      }
      if (opt.output_phased_gls) {
        out << ":" << fixed2(phased_gls[0]);
        for (int i = 0; i < V; i++)
          for (int j = 0; j < V; j++)
            if (i != 0 || j != 0) out << "," << fixed2(phased_gls[new_to_old[i] * V + new_to_old[j]]);
      }
    }
    if (opt.output_haplotype_data && (output_lflanks || output_rflanks)) {
      out << ":" << fixed2(std::exp(k3b.hap_log_unphased[s])) << ":" << fixed2(std::exp(k3b.hap_log_phased[s]));
      if (!g.haploid_) {
        if (output_lflanks) out << ":" << hap_to_lflank[hap_a] << "|" << hap_to_lflank[hap_b];
        if (output_rflanks) out << ":" << hap_to_rflank[hap_a] << "|" << hap_to_rflank[hap_b];
      } else {
        if (output_lflanks) out << ":" << hap_to_lflank[hap_a];
        if (output_rflanks) out << ":" << hap_to_rflank[hap_a];
      }
    }
    if (opt.output_filters) out << ":PASS";
  }
  g.vcf_record_ = out.str();
  g.vcf_pos_ = pos;
  if (!filter_reasons.empty()) {
    int total = 0;
    for (const auto& kv : filter_reasons) total += kv.second;
    std::ostringstream msg;
    msg << "Filtered " << total << " sample genotypes for the following reasons:\t";
    for (const auto& kv : filter_reasons) msg << kv.second << "=" << kv.first << "\t";
    g.log_ += msg.str() + "\n";
  }
}

}  // namespace

hipstr_status_t GenotyperBatch::write_vcf_records(const hipstr_vcf_loci_t* regions, const hipstr_vcf_options_t* options,
                                                  std::string& err) {
  if (!regions || !options) { err = "null argument"; return HIPSTR_ERR_BAD_ARG; }
  if (!ctx_) { err = "no device context"; return HIPSTR_ERR_NO_DEVICE; }
  const double t_begin = now_s();
  const double other0 = seconds[T_TRACE_DEVICE];
  // K3b over every genotyped locus
  std::vector<int> which;
  std::vector<int32_t> locus_sample_off{0}, n_haps, n_variants, hap_to_allele;
  std::vector<uint8_t> haploid;
  std::vector<double> post, sample_ll;
  std::vector<size_t> gl_off{0}, pgl_off{0};
  std::vector<int> sample_base;   // first row of the locus in locus_sample_names
  {
    int base = 0;
    for (size_t l = 0; l < loci.size(); l++) { sample_base.push_back(base); base += loci[l].num_samples_; }
  }
  for (size_t l = 0; l < loci.size(); l++) {
    SeqStutterGenotyper& g = loci[l];
    g.vcf_record_.clear();
    if (!g.succeeded()) continue;
    int rep = -1, n_rep = 0;
    for (size_t b = 0; b < g.hap_blocks_.size(); b++)
      if (g.hap_blocks_[b].period > 0) { if (rep < 0) rep = (int)b; n_rep++; }
    if (n_rep != 1) { err = "write_vcf_records handles one STR region per locus"; return HIPSTR_ERR_UNSUPPORTED; }
    which.push_back((int)l);
    const int V = g.hap_blocks_[rep].num_options();
    std::vector<int> h2a;
    g.haps_to_alleles(rep, h2a);
    hap_to_allele.insert(hap_to_allele.end(), h2a.begin(), h2a.end());
    n_haps.push_back(g.num_alleles_);
    n_variants.push_back(V);
    haploid.push_back(g.haploid_ ? 1 : 0);
    locus_sample_off.push_back(locus_sample_off.back() + g.num_samples_);
    post.insert(post.end(), g.log_sample_posteriors_.begin(), g.log_sample_posteriors_.end());
    sample_ll.insert(sample_ll.end(), g.sample_total_LLs_.begin(), g.sample_total_LLs_.end());
    gl_off.push_back(gl_off.back() + (size_t)g.num_samples_ * (g.haploid_ ? V : V * (V + 1) / 2));
    pgl_off.push_back(pgl_off.back() + (size_t)g.num_samples_ * (g.haploid_ ? V : V * V));
  }
  if (which.empty()) return HIPSTR_OK;
  const size_t S = locus_sample_off.back();
  std::vector<int32_t> best_hap(2 * S), best_gt(2 * S), pl(gl_off.back());
  std::vector<double> log_phased(S), log_unphased(S), hap_log_phased(S), hap_log_unphased(S), gl(gl_off.back()), phased_gl(pgl_off.back()),
      gl_diff(S);
  hipstr_status_t st = hipstr_extract_genotypes_host(ctx_, (int32_t)which.size(), locus_sample_off.data(), n_haps.data(), n_variants.data(),
                                                     hap_to_allele.data(), haploid.data(), post.data(), sample_ll.data(), best_hap.data(),
                                                     best_gt.data(), log_phased.data(), log_unphased.data(), hap_log_phased.data(),
                                                     hap_log_unphased.data(), gl.data(), phased_gl.data(), gl_diff.data(), pl.data());
  account_device_call();
  if (st != HIPSTR_OK) { err = std::string("hipstr_extract_genotypes_host: ") + hipstr_last_error(ctx_); return st; }
  // traces of the reads against the haplotype their strand assignment picks
  parallel_for(which.size(), [&](size_t k) { loci[which[k]].vcf_prepare(&best_hap[2 * (size_t)locus_sample_off[k]]); });
  st = run_traces(which, err);
  if (st != HIPSTR_OK) return st;
  // a view of every locus' chromosome (the C-ABI hands NUL-terminated sequences: one strlen per distinct chromosome, and no
  // copy -- a std::string parameter here copied the whole chromosome for every record)
  std::vector<std::string_view> chrom_of(which.size());
  {
    std::map<const char*, size_t> length_of;
    for (size_t k = 0; k < which.size(); k++) {
      const char* c = regions->chrom_seq[which[k]];
      auto it = length_of.find(c);
      if (it == length_of.end()) it = length_of.emplace(c, c ? std::strlen(c) : 0).first;
      chrom_of[k] = std::string_view(c ? c : "", it->second);
    }
  }
  parallel_for(which.size(), [&](size_t k) {
    const int l = which[k];
    SeqStutterGenotyper& g = loci[l];
    const size_t s0 = locus_sample_off[k];
    LocusGenotypes v;
    v.best_hap = &best_hap[2 * s0]; v.best_gt = &best_gt[2 * s0];
    v.log_phased = &log_phased[s0]; v.log_unphased = &log_unphased[s0];
    v.hap_log_phased = &hap_log_phased[s0]; v.hap_log_unphased = &hap_log_unphased[s0];
    v.gl = gl.data() + gl_off[k]; v.pl = pl.data() + gl_off[k]; v.phased_gl = phased_gl.data() + pgl_off[k];
    v.gl_diff = &gl_diff[s0];
    v.n_gl = g.num_samples_ ? (int)((gl_off[k + 1] - gl_off[k]) / g.num_samples_) : 0;
    v.n_phased_gl = g.num_samples_ ? (int)((pgl_off[k + 1] - pgl_off[k]) / g.num_samples_) : 0;
    std::vector<std::string> locus_names, out_names;
    for (int s = 0; s < g.num_samples_; s++) locus_names.push_back(regions->locus_sample_names[sample_base[l] + s]);
    for (int s = 0; s < regions->n_out_samples; s++) out_names.push_back(regions->out_sample_names[s]);
    format_record(g, v, regions->chrom[l], regions->name && regions->name[l] ? regions->name[l] : "", regions->region_start[l],
                  regions->region_stop[l], regions->period[l], chrom_of[k], locus_names, out_names, *options);
  });
  seconds[T_VCF] += (now_s() - t_begin) - (seconds[T_TRACE_DEVICE] - other0);
  return HIPSTR_OK;
}

}  // namespace hipstr

extern "C" {

double hipstr_allele_bias(int32_t hap_a_reads, int32_t hap_b_reads) { return hipstr::compute_allele_bias(hap_a_reads, hap_b_reads); }
double hipstr_fisher_two_sided(int32_t n11, int32_t n12, int32_t n21, int32_t n22) { return hipstr::fisher_two_sided(n11, n12, n21, n22); }
int32_t hipstr_extract_cigar(const char* cigar_type, const int32_t* cigar_len, int32_t n, int32_t cigar_start, int32_t region_start,
                             int32_t region_end, int32_t* bp_diff) {
  int d = 0;
  if (!cigar_type || !cigar_len || !bp_diff || n <= 0) return 0;
  const bool ok = hipstr::extract_cigar(cigar_type, cigar_len, n, cigar_start, region_start, region_end, d);
  *bp_diff = d;
  return ok ? 1 : 0;
}

void hipstr_vcf_default_options(hipstr_vcf_options_t* o) {   // genotyper.cpp:336-343
  if (!o) return;
  o->output_gls = 0; o->output_pls = 0; o->output_phased_gls = 0; o->output_allreads = 1; o->output_mallreads = 1;
  o->output_filters = 0; o->output_haplotype_data = 0;
  o->max_flank_indel_frac = 0.15;
}

hipstr_status_t hipstr_genotyper_write_vcf(hipstr_genotyper_t* g, const hipstr_vcf_loci_t* loci, const hipstr_vcf_options_t* options) {
  if (!g) return HIPSTR_ERR_BAD_ARG;
  hipstr_vcf_options_t def;
  hipstr_vcf_default_options(&def);
  return g->batch.write_vcf_records(loci, options ? options : &def, g->last_error);
}

int32_t hipstr_genotyper_locus_record(const hipstr_genotyper_t* g, int32_t locus, int32_t* pos, char* out, int32_t cap) {
  if (!g || locus < 0 || locus >= (int32_t)g->batch.loci.size()) return -1;
  const hipstr::SeqStutterGenotyper& s = g->batch.loci[locus];
  if (pos) *pos = s.vcf_pos_;
  if (!out || (int32_t)s.vcf_record_.size() + 1 > cap) return -(int32_t)s.vcf_record_.size() - 1;
  std::memcpy(out, s.vcf_record_.c_str(), s.vcf_record_.size() + 1);
  return (int32_t)s.vcf_record_.size();
}

hipstr_status_t hipstr_genotyper_emit_records(const hipstr_genotyper_t* g, const hipstr_vcf_loci_t* loci, hipstr_vcf_writer_t* w) {
  if (!g || !loci || !w) return HIPSTR_ERR_BAD_ARG;
  hipstr::VCFWriter* writer = reinterpret_cast<hipstr::VCFWriter*>(w);
  for (size_t l = 0; l < g->batch.loci.size(); l++) {
    const hipstr::SeqStutterGenotyper& s = g->batch.loci[l];
    if (s.vcf_record_.empty()) continue;
    if (!writer->add_vcf_record(loci->chrom[l], s.vcf_pos_, s.vcf_record_)) return HIPSTR_ERR_BAD_ARG;
  }
  return HIPSTR_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// The VCF header (Genotyper::get_vcf_header, src/genotyper.cpp:253-331): the field dictionary of the records that
// format_record writes, as a table.  The texts are part of the file format the reference emits.
// ---------------------------------------------------------------------------------------------
namespace {
struct FieldDef { const char* id; const char* number; const char* type; const char* text; };
const FieldDef kInfoFields[] = {
    {"INFRAME_PGEOM", "1", "Float", "Parameter for in-frame geometric step size distribution"},
    {"INFRAME_UP", "1", "Float", "Probability that stutter causes an in-frame increase in obs. STR size"},
    {"INFRAME_DOWN", "1", "Float", "Probability that stutter causes an in-frame decrease in obs. STR size"},
    {"OUTFRAME_PGEOM", "1", "Float", "Parameter for out-of-frame geometric step size distribution"},
    {"OUTFRAME_UP", "1", "Float", "Probability that stutter causes an out-of-frame increase in read's STR size"},
    {"OUTFRAME_DOWN", "1", "Float", "Probability that stutter causes an out-of-frame decrease in read's STR size"},
    {"BPDIFFS", "A", "Integer", "Base pair difference of each alternate allele from the reference allele"},
    {"START", "1", "Integer", "Inclusive start coodinate for the repetitive portion of the reference allele"},
    {"END", "1", "Integer", "Inclusive end coordinate for the repetitive portion of the reference allele"},
    {"PERIOD", "1", "Integer", "Length of STR motif"},
    {"AN", "1", "Integer", "Total number of alleles in called genotypes"},
    {"REFAC", "1", "Integer", "Reference allele count"},
    {"AC", "A", "Integer", "Alternate allele counts"},
    {"NSKIP", "1", "Integer", "Number of samples not genotyped due to various issues"},
    {"NFILT", "1", "Integer", "Number of samples whose genotypes were filtered due to various issues"},
    {"DP", "1", "Integer", "Total number of valid reads used to genotype all samples"},
    {"DSNP", "1", "Integer", "Total number of reads with SNP phasing information"},
    {"DSTUTTER", "1", "Integer", "Total number of reads with a stutter indel in the STR region"},
    {"DFLANKINDEL", "1", "Integer", "Total number of reads with an indel in the regions flanking the STR"},
};
const FieldDef kHaplotypeInfoFields[] = {
    {"LFLANKS", ".", "String", "Comma-separated sequence(s) of flank to the  left of the repeat. Only output if 1 or more non-ref  left flanks were detected"},
    {"RFLANKS", ".", "String", "Comma-separated sequence(s) of flank to the right of the repeat. Only output if 1 or more non-ref right flanks were detected"},
};
const char kBiasTail[] = "where 0 is no bias and more negative values are increasingly biased. For homozygous genotypes, this can be negative if the haplotypes are heterozygous";
const FieldDef kFormatFields[] = {
    {"GT", "1", "String", "Genotype"},
    {"GB", "1", "String", "Base pair differences of genotype from reference"},
    {"Q", "1", "Float", "Posterior probability of unphased genotype"},
    {"PQ", "1", "Float", "Posterior probability of phased genotype"},
    {"DP", "1", "Integer", "Number of valid reads used for sample's genotype"},
    {"DSNP", "1", "Integer", "Number of reads with SNP phasing information"},
    {"PSNP", "1", "String", "Number of reads with SNPs supporting each haploid genotype"},
    {"PDP", "1", "String", "Fractional reads supporting each haploid genotype"},
    {"GLDIFF", "1", "Float", "Difference in likelihood between the reported and next best genotypes"},
    {"DSTUTTER", "1", "Integer", "Number of reads with a stutter indel in the STR region"},
    {"DFLANKINDEL", "1", "Integer", "Number of reads with an indel in the regions flanking the STR"},
    {"AB", "1", "Float", "log10 of the allele bias pvalue, "},       // + kBiasTail
    {"FS", "1", "Float", "log10 of the strand bias pvalue from Fisher's exact test, "},
    {"DAB", "1", "Integer", "Number of reads used in the AB and FS calculations"},
};
const FieldDef kHaplotypeFormatFields[] = {
    {"HQ", "1", "Float", "Posterior probability of unphased haplotypes. Only output if 1 or more non-ref flanks were detected"},
    {"PHQ", "1", "Float", "Posterior probability of   phased haplotypes. Only output if 1 or more non-ref flanks were detected"},
    {"LFGT", "1", "String", "Genotype of  left flank with corresponding sequences reported in LFLANKS. Only output if 1 or more non-ref  left flanks were detected"},
    {"RFGT", "1", "String", "Genotype of right flank with corresponding sequences reported in RFLANKS. Only output if 1 or more non-ref right flanks were detected"},
};
void put_field(std::ostringstream& out, const char* kind, const FieldDef& f, const char* tail = "") {
  out << "##" << kind << "=<ID=" << f.id << ",Number=" << f.number << ",Type=" << f.type << ",Description=\"" << f.text << tail << "\">\n";
}
}  // namespace

extern "C" int64_t hipstr_vcf_header(const char* reference_path, const char* full_command, int32_t n_contigs, const char* const* contig_names,
                                     const int64_t* contig_lengths, int32_t n_samples, const char* const* sample_names,
                                     const hipstr_vcf_options_t* o, int64_t cap, char* out_text) {
  if (!reference_path || !full_command || n_contigs < 0 || n_samples < 0 || !o || (n_contigs > 0 && (!contig_names || !contig_lengths)) ||
      (n_samples > 0 && !sample_names))
    return 0;
  std::ostringstream out;
  out << "##fileformat=VCFv4.1\n##command=" << full_command << "\n##reference=" << reference_path << "\n";
  for (int c = 0; c < n_contigs; c++) out << "##contig=<ID=" << contig_names[c] << ",length=" << contig_lengths[c] << ">\n";
  for (const FieldDef& f : kInfoFields) put_field(out, "INFO", f);
  if (o->output_haplotype_data)
    for (const FieldDef& f : kHaplotypeInfoFields) put_field(out, "INFO", f);
  for (const FieldDef& f : kFormatFields) put_field(out, "FORMAT", f, (std::strcmp(f.id, "AB") == 0 || std::strcmp(f.id, "FS") == 0) ? kBiasTail : "");
  if (o->output_haplotype_data)
    for (const FieldDef& f : kHaplotypeFormatFields) put_field(out, "FORMAT", f);
  if (o->output_allreads) put_field(out, "FORMAT", FieldDef{"ALLREADS", "1", "String", "Base pair difference observed in each read's Needleman-Wunsch alignment"});
  if (o->output_mallreads)
    put_field(out, "FORMAT", FieldDef{"MALLREADS", "1", "String", "Maximum likelihood bp diff in each read based on haplotype alignments for reads that span the repeat region by at least 5 base pairs"});
  if (o->output_gls) put_field(out, "FORMAT", FieldDef{"GL", "G", "Float", "log10 genotype likelihoods"});
  if (o->output_pls) put_field(out, "FORMAT", FieldDef{"PL", "G", "Integer", "Phred-scaled genotype likelihoods"});
  if (o->output_phased_gls)
    put_field(out, "FORMAT", FieldDef{"PHASEDGL", ".", "Float", "log10 genotype likelihood for each phased genotype. Value for phased genotype X|Y is stored at a 0-based index of X*A + Y, where A is the number of alleles. Not applicable to haploid genotypes"});
  if (o->output_filters) put_field(out, "FORMAT", FieldDef{"FILTER", "1", "String", "Reason for filtering the current call, or PASS if the call was not filtered"});
  out << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
  for (int s = 0; s < n_samples; s++) out << "\t" << sample_names[s];
  out << "\n";
  const std::string text = out.str();
  if (!out_text || (int64_t)text.size() + 1 > cap) return -(int64_t)text.size() - 1;
  std::memcpy(out_text, text.c_str(), text.size() + 1);
  return (int64_t)text.size();
}
