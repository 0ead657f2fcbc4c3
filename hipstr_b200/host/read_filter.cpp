/* read_filter.cpp -- see read_filter.h. */
#include "read_filter.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <sstream>

namespace hipstr {

namespace {

/* tolower() of the "C" locale, which is what the reference's comparisons run under, without the library call per character */
inline char low(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + ('a' - 'A')) : c; }

/* number of leading characters of `pattern` equal (ignoring case) to text[at...] */
int prefix_matches(std::string_view pattern, std::string_view text, int at) {
  int n = 0;
  while (n < (int)pattern.size() && at + n < (int)text.size() && low(pattern[n]) == low(text[at + n])) n++;
  return n;
}
/* number of trailing characters of `pattern` equal (ignoring case) to text[...at], read backwards from `at` */
int suffix_matches(std::string_view pattern, std::string_view text, int at) {
  int n = 0;
  while (n < (int)pattern.size() && at - n >= 0 && low(pattern[pattern.size() - 1 - n]) == low(text[at - n])) n++;
  return n;
}

std::string reverse_complement(const std::string& s) {
  std::string out;
  for (auto it = s.rbegin(); it != s.rend(); ++it) {
    switch (*it) {
      case 'A': case 'a': out += 'T'; break;
      case 'C': case 'c': out += 'G'; break;
      case 'G': case 'g': out += 'C'; break;
      case 'T': case 't': out += 'A'; break;
      default: throw FilterError(std::string("Invalid character in pattern argument to reverse_complement(): ") + *it);
    }
  }
  return out;
}

std::vector<std::string> split(const std::string& s, char delim) {   // getline semantics: no trailing empty item
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, delim)) out.push_back(item);
  return out;
}

bool consumes_read_and_ref(char t) { return t == 'M' || t == '=' || t == 'X'; }

const double* correct_table() {   // BaseQuality(): log P(base correct) by quality index (base_quality.h:29-38)
  static double table[42];
  static bool ready = false;
  if (!ready) {
    table[0] = -100000;
    for (int i = 1; i <= 41; i++) table[i] = std::log(1.0 - std::pow(10.0, i / (-10.0)));
    ready = true;
  }
  return table;
}
const double* const kCorrect = correct_table();

}  // namespace

std::string cigar_string(const std::vector<std::pair<char, int32_t> >& cigar) {
  std::stringstream ss;
  for (const auto& op : cigar) ss << op.second << op.first;
  return ss.str();
}

// ---------------------------------------------------------------------------------------------
// trimming (bam_io.cpp:384-547)
// ---------------------------------------------------------------------------------------------
void trim_alignment(BamRecord& a, int32_t min_read_start, int32_t max_read_stop, char min_base_qual) {
  auto check = [](char t) {
    if (!(consumes_read_and_ref(t) || t == 'I' || t == 'S' || t == 'D' || t == 'H')) throw FilterError("Invalid CIGAR option encountered in TrimAlignment");
  };
  int ltrim = 0;
  int32_t start_pos = a.pos;
  size_t front = 0;
  while (start_pos < min_read_start && front < a.cigar.size()) {
    const char t = a.cigar[front].first;
    check(t);
    const bool has_base = t != 'D' && t != 'H';
    if (has_base && a.quals[ltrim] > min_base_qual) break;
    if (has_base) ltrim++;
    if (consumes_read_and_ref(t) || t == 'D') start_pos++;
    if (--a.cigar[front].second == 0) front++;
  }
  a.cigar.erase(a.cigar.begin(), a.cigar.begin() + front);
  int rtrim = 0;
  const int last = (int)a.quals.size() - 1;
  int32_t end_pos = a.end_pos;
  while (end_pos > max_read_stop && !a.cigar.empty()) {
    const char t = a.cigar.back().first;
    check(t);
    const bool has_base = t != 'D' && t != 'H';
    if (has_base && a.quals[last - rtrim] > min_base_qual) break;
    if (has_base) rtrim++;
    if (consumes_read_and_ref(t) || t == 'D') end_pos--;
    if (--a.cigar.back().second == 0) a.cigar.pop_back();
  }
  if (ltrim + rtrim > (int)a.bases.size()) throw FilterError("CIGAR string does not correspond to alignment bases");
  a.bases = a.bases.substr(ltrim, a.bases.size() - ltrim - rtrim);
  a.quals = a.quals.substr(ltrim, a.quals.size() - ltrim - rtrim);
  a.pos = start_pos;
  a.end_pos = end_pos;
}

void trim_num_bases(BamRecord& a, int left_trim, int right_trim) {
  if (left_trim + right_trim > (int)a.bases.size()) throw FilterError("TrimNumBases: more bases to trim than the read has");
  auto step = [](char t, int& remaining, int32_t& coord, int dir) {
    if (consumes_read_and_ref(t)) { remaining--; coord += dir; }
    else if (t == 'D') coord += dir;
    else if (t == 'I' || t == 'S') remaining--;
    else if (t != 'H') throw FilterError("Invalid CIGAR option encountered in TrimAlignment");
  };
  int rem_l = left_trim;
  int32_t start_pos = a.pos;
  size_t front = 0;
  while (front < a.cigar.size() && (rem_l > 0 || a.cigar[front].first == 'D')) {
    step(a.cigar[front].first, rem_l, start_pos, +1);
    if (--a.cigar[front].second == 0) front++;
  }
  a.cigar.erase(a.cigar.begin(), a.cigar.begin() + front);
  int rem_r = right_trim;
  int32_t end_pos = a.end_pos;
  while (!a.cigar.empty() && (rem_r > 0 || a.cigar.back().first == 'D')) {
    step(a.cigar.back().first, rem_r, end_pos, -1);
    if (--a.cigar.back().second == 0) a.cigar.pop_back();
  }
  if (rem_l != 0 || rem_r != 0) throw FilterError("TrimNumBases: CIGAR shorter than the bases to trim");
  a.bases = a.bases.substr(left_trim, a.bases.size() - left_trim - right_trim);
  a.quals = a.quals.substr(left_trim, a.quals.size() - left_trim - right_trim);
  a.pos = start_pos;
  a.end_pos = end_pos;
}

// ---------------------------------------------------------------------------------------------
// adapter trimming (adapter_trimmer.{h,cpp})
// ---------------------------------------------------------------------------------------------
namespace {
const int kMinOverlap = 5;            // AdapterTrimmer::MIN_OVERLAP
const double kMaxErrorRate = 0.15;    // AdapterTrimmer::MAX_ERROR_RATE
}  // namespace

AdapterTrimmer::AdapterTrimmer() {
  r1_fw_ = {"AGATCGGAAGAGCAC", "CTGTCTCTTATACAC"};   // TruSeq R1, Nextera R1 (adapter_trimmer.cpp:196-199)
  r2_fw_ = {"AGATCGGAAGAGCGT", "CTGTCTCTTATACAC"};
  for (const std::string& s : r1_fw_) r1_rc_.push_back(reverse_complement(s));
  for (const std::string& s : r2_fw_) r2_rc_.push_back(reverse_complement(s));
}

/* The adapter (reverse complemented) sits at the START of a reverse-strand read: find the right-most read index at which an
 * adapter ends with at most one mismatch (the part of the adapter hanging off the left end is free) and cut through it. */
int64_t AdapterTrimmer::trim_five_prime(BamRecord& a, const std::vector<std::string>& adapters) const {
  const std::string& bases = a.bases;
  const int read_length = (int)bases.size();
  int trim_index = -1;
  for (const std::string& adapter : adapters) {
    const int adapter_length = (int)adapter.size();
    const std::string head = bases.substr(0, std::min(read_length, adapter_length));
    for (int index = read_length - 1; index >= kMinOverlap - 1; --index) {
      const int max_match = std::min(adapter_length, index + 1);
      const int tail_run = suffix_matches(adapter, bases, index);
      bool valid = tail_run == max_match;
      if (!valid && 1.0 / max_match < kMaxErrorRate) {
        if (max_match < adapter_length) valid = tail_run + 1 + prefix_matches(head, adapter, adapter_length - max_match) == max_match;
        else valid = tail_run + 1 + prefix_matches(adapter, bases, index - adapter_length + 1) == adapter_length;
      }
      if (valid) {
        trim_index = std::max(trim_index, index);
        break;
      }
    }
  }
  if (trim_index >= 0) trim_num_bases(a, trim_index + 1, 0);
  return trim_index + 1;
}

/* Mirror image for forward-strand reads: the left-most index at which an adapter starts. */
int64_t AdapterTrimmer::trim_three_prime(BamRecord& a, const std::vector<std::string>& adapters) const {
  const std::string& bases = a.bases;
  const int read_length = (int)bases.size();
  int trim_index = read_length;
  for (const std::string& adapter : adapters) {
    const int adapter_length = (int)adapter.size();
    const std::string tail = bases.substr(bases.size() - std::min(read_length, adapter_length));
    for (int index = 0; index <= read_length - kMinOverlap; ++index) {
      const int max_match = std::min(adapter_length, read_length - index);
      const int head_run = prefix_matches(adapter, bases, index);
      bool valid = head_run == max_match;
      if (!valid && 1.0 / max_match < kMaxErrorRate) {
        if (max_match < adapter_length) valid = head_run + 1 + suffix_matches(tail, adapter, max_match - 1) == max_match;
        else valid = head_run + 1 + suffix_matches(adapter, bases, index + adapter_length - 1) == adapter_length;
      }
      if (valid) {
        trim_index = std::min(trim_index, index);
        break;
      }
    }
  }
  if (trim_index < read_length) trim_num_bases(a, 0, read_length - trim_index);
  return read_length - trim_index;
}

void AdapterTrimmer::trim_adapters(BamRecord& a) {
  if (!trim_ || a.length() == 0) return;
  if (a.first_mate() || !a.paired()) {
    const int64_t n = a.reverse() ? trim_five_prime(a, r1_rc_) : trim_three_prime(a, r1_fw_);
    r1_trimmed_bases += n;
    r1_trimmed_reads += n > 0;
    r1_total_reads++;
  } else if (a.second_mate()) {
    const int64_t n = a.reverse() ? trim_five_prime(a, r2_rc_) : trim_three_prime(a, r2_fw_);
    r2_trimmed_bases += n;
    r2_trimmed_reads += n > 0;
    r2_total_reads++;
  } else throw FilterError(a.name);
}

std::string AdapterTrimmer::stats_message() const {
  std::stringstream msg;
  msg << std::setprecision(2) << "Adapter trimming removed\n\t" << r1_trimmed_bases << " likely adapter bases from " << r1_trimmed_reads << "/"
      << r1_total_reads << " R1 reads (" << (r1_total_reads == 0 ? 0 : 100.0 * r1_trimmed_reads / r1_total_reads) << "%)\n\t" << r2_trimmed_bases
      << " likely adapter bases from " << r2_trimmed_reads << "/" << r2_total_reads << " R2 reads ("
      << (r2_total_reads == 0 ? 0 : 100.0 * r2_trimmed_reads / r2_total_reads) << "%)";
  return msg.str();
}

// ---------------------------------------------------------------------------------------------
// single-read filters (alignment_filters.cpp)
// ---------------------------------------------------------------------------------------------
namespace filters {

namespace {
template <class It>
int dist_to_indel(It it, It end) {
  if (it != end && it->first == 'H') ++it;
  if (it != end && it->first == 'S') ++it;
  int dist = 0;
  for (; it != end; ++it) {
    const char t = it->first;
    if (t == 'M') dist += it->second;
    else if (t == 'I' || t == 'D') return dist;
    else if (t == 'S' || t == 'H') return -1;
    else throw FilterError(std::string("Invalid CIGAR char") + t);
  }
  return -1;
}
}  // namespace

std::pair<int, int> end_dist_to_indel(const BamRecord& a) {
  return std::make_pair(dist_to_indel(a.cigar.begin(), a.cigar.end()), dist_to_indel(a.cigar.rbegin(), a.cigar.rend()));
}

std::pair<int, int> num_end_matches(const BamRecord& a, std::string_view ref, int ref_seq_start) {
  if (a.pos < ref_seq_start) return std::make_pair(-1, -1);
  size_t read_index = 0, ref_index = (size_t)(a.pos - ref_seq_start);
  auto it = a.cigar.begin();
  const auto end = a.cigar.end();
  bool in_head = true;   // no mismatch / indel seen yet
  int run = 0, head = 0;
  auto interrupt = [&]() { if (in_head) head = run; in_head = false; run = 0; };
  if (it != end && it->first == 'H') ++it;
  if (it != end && it->first == 'S') { read_index += it->second; ++it; }
  for (; it != end && ref_index < ref.size() && read_index < a.bases.size(); ++it) {
    const char t = it->first;
    const size_t len = (size_t)it->second;
    if (t == 'M') {
      if (ref_index + len > ref.size()) return std::make_pair(-1, -1);
      if (read_index + len > a.bases.size()) throw FilterError("Nucleotides for aligned read don't correspond to the CIGAR string");
      for (size_t k = 0; k < len; k++, read_index++, ref_index++) {
        if (low(ref[ref_index]) == low(a.bases[read_index])) run++;
        else interrupt();
      }
    } else if (t == 'I') { interrupt(); read_index += len; }
    else if (t == 'D') { interrupt(); ref_index += len; }
    else if (t == 'S' || t == 'H') break;
    else throw FilterError(std::string("Invalid CIGAR char") + t);
  }
  if (it != end && it->first == 'S') { read_index += it->second; ++it; }
  if (it != end && it->first == 'H') ++it;
  if (it != end) {
    if (ref_index >= ref.size()) return std::make_pair(-1, -1);
    throw FilterError("Improperly formatted CIGAR string");
  }
  if (read_index != a.bases.size()) {
    if (ref_index >= ref.size()) return std::make_pair(-1, -1);
    throw FilterError("CIGAR string does not correspond to alignment bases");
  }
  return in_head ? std::make_pair(run, run) : std::make_pair(head, run);
}

bool has_largest_end_matches(const BamRecord& a, std::string_view ref, int ref_seq_start, int max_external, int max_internal) {
  // GetUnclippedInfo: the aligned part of the read and its first / last reference coordinate
  int32_t start = a.pos, last = a.pos - 1;
  bool leading = true;
  int first_base = 0, n_bases = 0;
  for (const auto& op : a.cigar) {
    switch (op.first) {
      case 'D': last += op.second; leading = false; break;
      case 'H': break;
      case 'S': if (leading) first_base += op.second; break;
      case 'M': last += op.second; n_bases += op.second; leading = false; break;
      case 'I': n_bases += op.second; leading = false; break;
      default: throw FilterError(std::string("Invalid CIGAR char ") + op.first);
    }
  }
  if ((size_t)first_base > a.bases.size()) throw FilterError("CIGAR string does not correspond to alignment bases");
  const std::string bases = a.bases.substr(first_base, n_bases);
  const int ref_len = (int)ref.size();
  // the read start must match longer at its own position than anywhere else in the window
  if (start >= ref_seq_start && start < ref_seq_start + ref_len) {
    const int at = start - ref_seq_start;
    const int lo = std::max(0, at - max_external), hi = std::min(ref_len - 1, at + max_internal);
    const int own = prefix_matches(bases, ref, at);
    for (int i = lo; i <= hi; i++)
      if (i != at && prefix_matches(bases, ref, i) >= own) return false;
  }
  if (last >= ref_seq_start && last < ref_seq_start + ref_len) {
    const int at = last - ref_seq_start;
    const int lo = std::max(0, at - max_internal), hi = std::min(ref_len - 1, at + max_external);
    const int own = suffix_matches(bases, ref, at);
    for (int i = lo; i <= hi; i++)
      if (i != at && suffix_matches(bases, ref, i) >= own) return false;
  }
  return true;
}

double sum_log_prob_correct(const std::string& quals) {
  double sum = 0.0;
  for (char q : quals) sum += kCorrect[q < '!' ? 0 : (q > 'J' ? 41 : q - '!')];
  return sum;
}

}  // namespace filters

// ---------------------------------------------------------------------------------------------
// alternate mappings and mate pairing (bam_processor.cpp:59-157)
// ---------------------------------------------------------------------------------------------
namespace {
const std::string& ref_name(int32_t id, const std::vector<std::string>& names) {
  static const std::string star = "*";
  if (id == -1) return star;
  if (id < 0 || id >= (int32_t)names.size()) throw FilterError("Invalid reference ID provided to ref_name() function");
  return names[id];
}
bool ends_with(const std::string& s, const std::string& suffix) { return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0; }
bool starts_with(const std::string& s, const std::string& prefix) { return s.size() >= prefix.size() && s.compare(0, prefix.size(), prefix) == 0; }
}  // namespace

void ReadFilter::extract_mappings(const BamRecord& a, const std::vector<std::string>& ref_names,
                                  std::vector<std::pair<std::string, int32_t> >& out) const {
  const std::string& chrom = ref_name(a.ref_id, ref_names);
  if (chrom == "*" || a.cigar.empty()) return;
  out.emplace_back(chrom, a.pos);
  std::string own_cigar;
  for (int which = 0; which < 2; which++) {
    if (!(which == 0 ? a.has_xa : a.has_sa)) continue;
    for (const std::string& alt : split(which == 0 ? a.xa : a.sa, ';')) {
      const std::vector<std::string> tokens = split(alt, ',');
      if (tokens.size() < 2) throw FilterError("Failed to extract XA or SA tag from BAM alignment");
      char* stop = nullptr;
      const long parsed = std::strtol(tokens[1].c_str(), &stop, 10);
      if (stop == tokens[1].c_str()) throw FilterError("Failed to extract XA or SA tag from BAM alignment");
      const int32_t pos = (int32_t)std::labs(parsed);
      if (tokens[0] == out[0].first && std::abs(pos - out[0].second) <= 200) continue;
      // GRCh38: a hit on an alt contig of the same chromosome with the same CIGAR is not a second mapping
      if (which == 0 && ends_with(tokens[0], "_alt") && starts_with(tokens[0], out[0].first + "_")) {
        if (own_cigar.empty()) own_cigar = cigar_string(a.cigar);
        if (tokens.size() > 2 && tokens[2] == own_cigar) continue;
      }
      out.emplace_back(tokens[0], pos);
    }
  }
}

void ReadFilter::valid_pairings(const BamRecord& a1, const BamRecord& a2, const std::vector<std::string>& ref_names,
                                std::vector<std::pair<std::string, int32_t> >& p1, std::vector<std::pair<std::string, int32_t> >& p2) const {
  if (a1.ref_id == -1 || a2.ref_id == -1) return;
  // BWA-MEM may omit XA when there are too many alternate hits: without XA, a read whose best score is within 10 of
  // its suboptimal score cannot vouch for the pair (the mate is asked first, then the read itself)
  if (!a2.has_xa) {
    if (a2.has_as && a2.has_xs && a2.as - a2.xs < 10) return;
  } else if (!a1.has_xa) {
    if (a1.has_as && a1.has_xs && a1.as - a1.xs < 10) return;
  }
  std::vector<std::pair<std::string, int32_t> > m1, m2;
  extract_mappings(a1, ref_names, m1);
  extract_mappings(a2, ref_names, m2);
  std::sort(m1.begin(), m1.end());
  std::sort(m2.begin(), m2.end());
  size_t min_j = 0;
  for (size_t i = 0; i < m1.size(); i++) {
    for (size_t j = min_j; j < m2.size(); j++) {
      const int cmp = m1[i].first.compare(m2[j].first);
      if (cmp < 0) break;
      if (cmp > 0) { min_j = j + 1; continue; }
      if (std::abs(m1[i].second - m2[j].second) < options.max_mate_dist) { p1.push_back(m1[i]); p2.push_back(m2[j]); }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// read_and_filter_reads (bam_processor.cpp:173-474)
// ---------------------------------------------------------------------------------------------
void ReadFilter::run(const std::vector<BamRecord>& records, const std::vector<std::string>& ref_names, const std::vector<std::string>& file_names,
                     std::string_view chrom_seq, const std::vector<std::pair<int32_t, int32_t> >& regions,
                     const std::map<std::string, std::string>& rg_to_sample, FilteredReads& out) {
  if (regions.empty()) throw FilterError("no region given");
  int32_t group_start = INT_MAX, group_stop = INT_MIN;
  for (const auto& r : regions) { group_start = std::min(group_start, r.first); group_stop = std::max(group_stop, r.second); }
  adapter_trimmer.set_enabled(options.trim_adapters);
  FilterCounts& n = out.counts;
  n = FilterCounts();
  std::vector<BamRecord> paired_strs, mate_alns, unpaired_strs;
  std::map<std::string, BamRecord> potential_strs, potential_mates;
  int32_t prev_file = -1, file_number = 0;
  std::string file_label = "0_";
  auto overlaps = [&](const BamRecord& a) { return a.pos < group_stop && a.end_pos >= group_start; };
  auto trimmed_name = [](const BamRecord& a) {   // "name/1" -> "name"
    std::string s = a.name;
    if (s.size() > 2 && s[s.size() - 2] == '/') s.resize(s.size() - 2);
    return s;
  };
  auto unique_pair = [&](const BamRecord& str_read, const BamRecord& mate) {
    std::vector<std::pair<std::string, int32_t> > p1, p2;
    valid_pairings(str_read, mate, ref_names, p1, p2);
    return p1.size() == 1 && p1[0].second == str_read.pos;
  };

  for (const BamRecord& record : records) {
    if (record.paired() && !record.first_mate() && !record.second_mate()) continue;
    // neither the read nor (judging by its mate's position) its mate can reach the STR
    if (record.pos > group_stop || record.end_pos < group_start) {
      if (!record.paired() || record.mate_pos == record.pos) continue;
      if (record.mate_pos > group_stop) continue;
      if (record.mate_pos + record.length() + 100 < group_start) continue;
    }
    if ((int64_t)paired_strs.size() > options.max_total_reads) { n.too_many_reads = true; break; }
    if (!record.mapped() || record.pos == 0 || record.cigar.empty() || record.length() == 0) continue;

    BamRecord a = record;
    if (overlaps(a)) {
      if (options.base_qual_trim > ' ') {
        if (a.cigar.front().first == 'H' || a.cigar.back().first == 'H') {   // trimming needs the clipped bases
          n.read_count++;
          n.hard_clip++;
          continue;
        }
        const int32_t length = a.length();
        trim_low_quality_ends(a, options.base_qual_trim);
        if (overlaps(a) && (a.length() == 0 || a.length() < length / 2)) continue;
      }
      adapter_trimmer.trim_adapters(a);
      if (a.cigar.empty() || a.length() == 0) continue;
    }

    if (prev_file != a.file) {   // a new file: mates seen so far cannot pair with its reads
      prev_file = a.file;
      potential_mates.clear();
      file_label = std::to_string(++file_number) + "_";
    }
    const std::string key = file_label + trimmed_name(a);

    if (overlaps(a)) {
      n.read_count++;
      bool pass_one = false;
      if (a.bases.find('N') != std::string::npos) n.read_has_n++;
      else if (filters::sum_log_prob_correct(a.quals) < options.min_sum_qual_log_prob) n.low_qual_score++;
      else pass_one = true;

      if (!pass_one) {
        potential_mates.insert(std::make_pair(key, a));
        continue;
      }
      // second set of filters, per STR of the group: may the read be used to generate candidate haplotypes?
      std::string pass_two(regions.size(), '0');
      for (size_t r = 0; r < regions.size(); r++) {
        if (options.min_flank > 0 && (a.pos > regions[r].first - options.min_flank || a.end_pos < regions[r].second + options.min_flank)) continue;
        bool ok = true;
        if (options.maximal_end_match_window > 0)
          ok = filters::has_largest_end_matches(a, chrom_seq, 0, options.maximal_end_match_window, options.maximal_end_match_window);
        if (ok && options.min_read_end_match > 0) {
          const std::pair<int, int> m = filters::num_end_matches(a, chrom_seq, 0);
          ok = m.first >= options.min_read_end_match && m.second >= options.min_read_end_match;
        }
        if (ok && options.min_bp_before_indel > 0) {
          const std::pair<int, int> d = filters::end_dist_to_indel(a);
          ok = !((d.first != -1 && d.first < options.min_bp_before_indel) || (d.second != -1 && d.second < options.min_bp_before_indel));
        }
        if (!ok) { pass_two.assign(regions.size(), '0'); break; }
        pass_two[r] = '1';
      }
      a.passes = pass_two;

      auto mate = potential_mates.find(key);
      if (mate != potential_mates.end()) {
        if (a.first_mate() == mate->second.first_mate()) {   // same end seen twice: keep the new one as an STR candidate
          potential_mates.erase(mate);
          potential_strs.insert(std::make_pair(key, a));
          continue;
        }
        if (unique_pair(a, mate->second)) { paired_strs.push_back(a); mate_alns.push_back(mate->second); }
        else n.unique_mapping++;
        potential_mates.erase(mate);
        continue;
      }
      auto other = potential_strs.find(key);
      if (other == potential_strs.end()) {
        potential_strs.insert(std::make_pair(key, a));
        continue;
      }
      if (a.first_mate() == other->second.first_mate()) { n.read_count--; continue; }
      if (unique_pair(a, other->second)) {   // both mates overlap the STR: each is genotyped with the other as its mate
        paired_strs.push_back(a); mate_alns.push_back(other->second);
        paired_strs.push_back(other->second); mate_alns.push_back(a);
      } else n.unique_mapping += 2;
      potential_strs.erase(other);
    } else {
      auto str_read = potential_strs.find(key);
      if (str_read != potential_strs.end()) {
        if (a.first_mate() == str_read->second.first_mate()) continue;
        if (unique_pair(str_read->second, a)) { paired_strs.push_back(str_read->second); mate_alns.push_back(a); }
        else n.unique_mapping++;
        potential_strs.erase(str_read);
        continue;
      }
      auto seen = potential_mates.find(key);
      if (seen == potential_mates.end()) potential_mates.insert(std::make_pair(key, a));
      else if (a.first_mate() != seen->second.first_mate()) potential_mates.erase(seen);
    }
  }

  for (const auto& kv : potential_strs) {   // STR reads whose mate never showed up, in key order
    if (kv.second.has_xa) n.unique_mapping++;
    else if (options.require_paired_reads) n.unpaired_filtered++;
    else unpaired_strs.push_back(kv.second);
  }

  // by read group, walking each list from its back (so every per-group list is in reverse order of discovery)
  out.rg_names.clear(); out.paired.clear(); out.mates.clear(); out.unpaired.clear();
  std::map<std::string, int> group_of;
  for (int type = 0; type < 2; type++) {
    std::vector<BamRecord>& src = type == 0 ? paired_strs : unpaired_strs;
    while (!src.empty()) {
      const BamRecord& a = src.back();
      if (!a.has_rg) throw FilterError("Failed to retrieve BAM alignment's RG tag");
      if (a.file < 0 || a.file >= (int32_t)file_names.size()) throw FilterError("alignment from an unknown file");
      auto sample = rg_to_sample.find(file_names[a.file] + a.rg);
      if (sample == rg_to_sample.end()) throw FilterError("No sample found for read group " + a.rg + " in BAM file headers");
      auto g = group_of.find(sample->second);
      if (g == group_of.end()) {
        g = group_of.insert(std::make_pair(sample->second, (int)out.rg_names.size())).first;
        out.rg_names.push_back(sample->second);
        out.paired.emplace_back(); out.mates.emplace_back(); out.unpaired.emplace_back();
      }
      if (type == 0) {
        out.paired[g->second].push_back(a);
        out.mates[g->second].push_back(mate_alns.back());
        mate_alns.pop_back();
      } else out.unpaired[g->second].push_back(a);
      src.pop_back();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// remove_pcr_duplicates (pcr_duplicates.{h,cpp})
// ---------------------------------------------------------------------------------------------
int32_t ReadFilter::remove_pcr_duplicates(const std::map<std::string, std::string>& rg_to_library, const std::vector<std::string>& file_names,
                                          FilteredReads& reads) {
  struct Pair {
    std::string library, name;
    int32_t min_start, max_start;   // min_start == -1: single-ended
    int index;                      // >= 0 paired[index]; < 0 unpaired[-index - 1]
    bool duplicate_of(const Pair& o) const { return library == o.library && min_start == o.min_start && max_start == o.max_start; }
    bool operator<(const Pair& o) const {
      const int c = library.compare(o.library);
      if (c != 0) return c < 0;
      if (min_start != o.min_start) return min_start < o.min_start;
      if (max_start != o.max_start) return max_start < o.max_start;
      return name.compare(o.name) < 0;
    }
  };
  auto library_of = [&](const BamRecord& a) {
    if (!a.has_rg) throw FilterError("Failed to retrieve BAM alignment's RG tag");
    auto it = rg_to_library.find(file_names.at(a.file) + a.rg);
    if (it == rg_to_library.end()) throw FilterError("No library found for read group " + a.rg + " in BAM file headers");
    return it->second;
  };
  int32_t dup_count = 0;
  for (size_t g = 0; g < reads.paired.size(); g++) {
    std::vector<BamRecord> paired, mates, unpaired;
    paired.swap(reads.paired[g]); mates.swap(reads.mates[g]); unpaired.swap(reads.unpaired[g]);
    std::vector<Pair> pairs;
    for (size_t j = 0; j < paired.size(); j++) {
      if (paired[j].name != mates[j].name) throw FilterError("mates with different names");
      pairs.push_back(Pair{library_of(paired[j]), paired[j].name, std::min(paired[j].pos, mates[j].pos), std::max(paired[j].pos, mates[j].pos), (int)j});
    }
    for (size_t j = 0; j < unpaired.size(); j++) pairs.push_back(Pair{library_of(unpaired[j]), unpaired[j].name, -1, unpaired[j].pos, -(int)j - 1});
    // the same comparator on the same initial order as the reference's std::sort of ReadPair objects: ties (a pair and
    // its mirror image when both mates overlap the STR) end up in the same places
    std::sort(pairs.begin(), pairs.end());
    if (pairs.empty()) continue;
    auto str_read = [&](const Pair& p) -> const BamRecord& { return p.index >= 0 ? paired[p.index] : unpaired[-p.index - 1]; };
    auto keep = [&](const Pair& p, bool include_rev) {
      if (p.min_start == -1) { reads.unpaired[g].push_back(unpaired[-p.index - 1]); return; }
      reads.paired[g].push_back(paired[p.index]);
      reads.mates[g].push_back(mates[p.index]);
      if (include_rev) {   // the mirror image of this pair was counted as its duplicate: restore it
        dup_count--;
        reads.paired[g].push_back(mates[p.index]);
        reads.mates[g].push_back(paired[p.index]);
      }
    };
    bool include_rev = false;
    size_t best = 0;
    for (size_t j = 1; j < pairs.size(); j++) {
      if (pairs[j].duplicate_of(pairs[best])) {
        dup_count++;
        if (filters::sum_log_prob_correct(str_read(pairs[j]).quals) > filters::sum_log_prob_correct(str_read(pairs[best]).quals)) {
          best = j;
          include_rev = pairs[best].name == pairs[j - 1].name;
        } else if (j == best + 1) include_rev |= pairs[best].name == pairs[j].name;
      } else {
        keep(pairs[best], include_rev);
        best = j;
        include_rev = false;
      }
    }
    keep(pairs[best], include_rev);
  }
  reads.counts.pcr_duplicates = dup_count;
  return dup_count;
}

}  // namespace hipstr
