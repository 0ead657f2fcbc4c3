/* haplotype_generator.cpp -- see haplotype_generator.h. */
#include "haplotype_generator.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <map>

namespace hipstr {

namespace {
const double kMinFracReads = 0.05, kMinFracSamples = 0.05, kMinFracStrongSample = 0.2, kMinReadsStrongSample = 2, kMinStrongSamples = 1;
const int32_t kLeftPad = 5, kRightPad = 5, kMinBlockSpacing = 10, kRefFlankLen = 35;   // HaplotypeGenerator.h:53-62

std::string upper(std::string s) {
  for (char& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}
bool by_length_then_sequence(const std::string& a, const std::string& b) {
  return a.size() != b.size() ? a.size() < b.size() : a.compare(b) < 0;
}
}  // namespace

bool HaplotypeGenerator::extract_sequence(const ReadView& aln, int32_t region_start, int32_t region_end, std::string& seq) {
  if (aln.start >= region_start || aln.stop <= region_end) return false;
  std::string& out = seq;   // built in place: the caller reuses one buffer for all its reads
  out.clear();
  // read bases are ASCII: upper-casing without the locale-aware call per character
  auto done = [&] { for (char& ch : out) ch = (ch >= 'a' && ch <= 'z') ? (char)(ch - 32) : ch; return true; };
  int32_t pos = aln.start;   // reference coordinate of the next unconsumed base of the current element
  int read_at = 0;           // next unconsumed read base
  for (int c = 0; c < aln.n_cigar; c++) {
    const char type = aln.cigar_type[c];
    const int len = aln.cigar_len[c];
    int used = 0;
    while (used < len) {
      if (pos > region_end) return done();
      if (pos == region_end) {
        if (type != 'I') return done();
        out.append(aln.bases + read_at, len);   // an insertion flush with the region's end still belongs to it
        read_at += len;
        used = len;
      } else if (pos >= region_start) {
        int n = std::min(region_end - pos, len - used);
        if (type == 'I') { n = len; out.append(aln.bases + read_at, n); read_at += n; }
        else if (type == '=' || type == 'X') { out.append(aln.bases + read_at, n); read_at += n; pos += n; }
        else if (type == 'D') pos += n;
        else return false;   // the reference dies on any other CIGAR operation
        used += n;
      } else {               // still left of the region
        int n;
        if (type == 'I') { n = len - used; read_at += n; }
        else {
          n = std::min(region_start - pos, len - used);
          pos += n;
          if (type != 'D') read_at += n;
        }
        used += n;
      }
    }
  }
  return false;   // unreachable for a spanning read (the reference dies with a logical error)
}

void HaplotypeGenerator::trim(int ideal_min_length, int32_t& region_start, int32_t& region_end, std::vector<std::string>& sequences) const {
  int min_len = INT_MAX;
  for (const std::string& s : sequences) min_len = std::min(min_len, (int)s.size());
  if (min_len <= ideal_min_length) return;
  // how far every sequence agrees from the left / from the right
  int max_left = 0, max_right = 0;
  while (max_left < min_len - ideal_min_length) {
    bool same = true;
    for (size_t j = 1; j < sequences.size() && same; j++) same = sequences[j][max_left] == sequences[j - 1][max_left];
    if (!same) break;
    max_left++;
  }
  while (max_right < min_len - ideal_min_length) {
    const char c = sequences[0][sequences[0].size() - 1 - max_right];
    bool same = true;
    for (size_t j = 1; j < sequences.size() && same; j++) same = sequences[j][sequences[j].size() - 1 - max_right] == c;
    if (!same) break;
    max_right++;
  }
  max_left = std::min(kLeftPad, max_left);     // never trim into the repeat itself
  max_right = std::min(kRightPad, max_right);
  max_left = std::max(0, std::min(min_len - kRightPad, max_left));
  max_right = std::max(0, std::min(min_len - kLeftPad, max_right));
  // clip as much as allowed, as evenly as possible
  int left, right;
  if (min_len - 2 * std::min(max_left, max_right) <= ideal_min_length) {
    left = right = std::min(max_left, max_right);
    while (min_len - left - right < ideal_min_length) {
      if (left > right) left--;
      else right--;
    }
  } else if (max_left > max_right) {
    right = max_right;
    left = std::min(max_left, min_len - ideal_min_length - max_right);
  } else {
    left = max_left;
    right = std::min(max_right, min_len - ideal_min_length - max_left);
  }
  for (std::string& s : sequences) s = s.substr(left, s.size() - left - right);
  region_start += left;
  region_end -= right;
}

void HaplotypeGenerator::gen_candidate_seqs(const std::string& ref_seq, int ideal_min_length,
                                            const std::vector<std::vector<ReadView> >& alignments, int32_t& region_start,
                                            int32_t& region_end, std::vector<std::string>& sequences) const {
  std::map<std::string, double> sample_counts;   // sum over samples of the fraction of the sample's reads
  std::map<std::string, int> read_counts, must_inc;
  int tot_reads = 0, tot_samples = 0;
  std::vector<std::pair<std::string, int> > counts;
  std::string sub;
  for (const std::vector<ReadView>& sample : alignments) {
    int samp_reads = 0;
    counts.clear();   // a sample carries a handful of distinct sequences: a flat list beats a tree of strings
    for (const ReadView& aln : sample) {
      if (extract_sequence(aln, region_start, region_end, sub)) {
        size_t k = 0;
        while (k < counts.size() && counts[k].first != sub) k++;
        if (k == counts.size()) counts.emplace_back(sub, 0);
        counts[k].second++;
        tot_reads++;
        samp_reads++;
      }
    }
    for (const auto& kv : counts) {   // alleles one sample supports strongly on its own
      read_counts[kv.first] += kv.second;
      if (kv.second >= kMinReadsStrongSample && kv.second >= kMinFracStrongSample * samp_reads) must_inc[kv.first]++;
      sample_counts[kv.first] += kv.second * 1.0 / samp_reads;
    }
    if (samp_reads > 0) tot_samples++;
  }
  int ref_index = -1;
  for (const auto& kv : must_inc)
    if (kv.second >= kMinStrongSamples) {
      sample_counts.erase(kv.first);
      read_counts.erase(kv.first);
      sequences.push_back(kv.first);
      if (kv.first == ref_seq) ref_index = (int)sequences.size() - 1;
    }
  for (const auto& kv : sample_counts)   // alleles with enough support across the population
    if (kv.second > kMinFracSamples * tot_samples || read_counts[kv.first] > kMinFracReads * tot_reads) {
      sequences.push_back(kv.first);
      if (ref_index == -1 && kv.first == ref_seq) ref_index = (int)sequences.size() - 1;
    }
  if (ref_index == -1) sequences.insert(sequences.begin(), ref_seq);
  else { sequences[ref_index] = sequences[0]; sequences[0] = ref_seq; }
  std::sort(sequences.begin() + 1, sequences.end(), by_length_then_sequence);
  trim(ideal_min_length, region_start, region_end, sequences);
}

bool HaplotypeGenerator::add_haplotype_block(int32_t reg_start, int32_t reg_stop, int32_t period, std::string_view chrom_seq,
                                             const std::vector<std::vector<ReadView> >& alignments, const double* stutter) {
  if (reg_start < kRefFlankLen + kLeftPad || reg_stop + kRefFlankLen + kRightPad > (int64_t)chrom_seq.size()) {
    failure_msg_ = "Haplotype blocks are too near to the chromosome ends";
    return false;
  }
  int32_t min_start = INT_MAX, max_stop = INT_MIN;
  for (const auto& sample : alignments)
    for (const ReadView& a : sample) { min_start = std::min(min_start, a.start); max_stop = std::max(max_stop, a.stop); }
  int32_t region_start = reg_start - kLeftPad, region_end = reg_stop + kRightPad;
  const std::string ref_seq = upper(std::string(chrom_seq.substr(region_start, region_end - region_start)));
  // With no alignment at all (every read of the locus failed the haplotype-generation filters) the reference's bounds stay
  // INT_MAX / INT_MIN and its "+ 5" / "- 5" wrap around, so the test passes and the block is built from the reference
  // allele alone; the wrap-around is reproduced here with unsigned arithmetic.
  const int32_t lo = (int32_t)((uint32_t)min_start + 5u), hi = (int32_t)((uint32_t)max_stop - 5u);
  if (lo >= region_start || hi <= region_end) {
    failure_msg_ = "No spanning alignments";
    return false;
  }
  std::vector<std::string> sequences;
  gen_candidate_seqs(ref_seq, 3 * period, alignments, region_start, region_end, sequences);
  if (!hap_blocks_.empty() && region_start < hap_blocks_.back().end + kMinBlockSpacing) {
    failure_msg_ = "Haplotype blocks are too near to one another";
    return false;
  }
  HapBlock block;
  block.start = region_start;
  block.end = region_end;
  block.period = period;
  std::memcpy(block.stutter, stutter, sizeof(block.stutter));
  block.seqs = sequences;
  hap_blocks_.push_back(block);
  return true;
}

bool HaplotypeGenerator::add_vcf_haplotype_block(int32_t pos, int32_t period, std::string_view chrom_seq,
                                                 const std::vector<std::string>& vcf_alleles, const double* stutter) {
  if (vcf_alleles.empty()) { failure_msg_ = "no alleles in the reference VCF record"; return false; }
  const int32_t region_start = pos, region_end = pos + (int32_t)vcf_alleles[0].size();
  if (region_start < kRefFlankLen || region_end + kRefFlankLen >= (int64_t)chrom_seq.size()) {
    failure_msg_ = "Haplotype blocks are too near to the chromosome ends";
    return false;
  }
  if (upper(vcf_alleles[0]) != upper(std::string(chrom_seq.substr(region_start, region_end - region_start)))) {
    failure_msg_ = "the reference allele of the VCF record does not match the chromosome sequence";   // an assert in the reference
    return false;
  }
  if (!hap_blocks_.empty() && region_start < hap_blocks_.back().end + kMinBlockSpacing) {
    failure_msg_ = "Haplotype blocks are too near to one another";
    return false;
  }
  HapBlock block;
  block.start = region_start;
  block.end = region_end;
  block.period = period;
  std::memcpy(block.stutter, stutter, sizeof(block.stutter));
  for (const std::string& a : vcf_alleles) block.seqs.push_back(upper(a));
  hap_blocks_.push_back(block);
  return true;
}

bool HaplotypeGenerator::fuse_haplotype_blocks(std::string_view chrom_seq) {
  if (hap_blocks_.empty()) { failure_msg_ = "no haplotype blocks were added"; return false; }
  // flanks of at most kRefFlankLen bp, at least 10 bp, no longer than the reads reach
  const int32_t min_start = std::min(hap_blocks_.front().start - 10, std::max(hap_blocks_.front().start - kRefFlankLen, min_aln_start_));
  const int32_t max_stop = std::max(hap_blocks_.back().end + 10, std::min(hap_blocks_.back().end + kRefFlankLen, max_aln_stop_));
  std::vector<HapBlock> fused;
  int32_t start = min_start;
  auto flank = [&](int32_t from, int32_t to) {
    HapBlock b;
    b.start = from;
    b.end = to;
    b.seqs.push_back(upper(std::string(chrom_seq.substr(from, to - from))));
    return b;
  };
  for (const HapBlock& b : hap_blocks_) {
    fused.push_back(flank(start, b.start));
    fused.push_back(b);
    start = b.end;
  }
  fused.push_back(flank(start, max_stop));
  hap_blocks_.swap(fused);
  return true;
}

}  // namespace hipstr
