/*
 * read_filter.h -- from the alignments of a region to the per-sample read lists the genotyper takes: host-side
 * restatement of BamProcessor::read_and_filter_reads (src/bam_processor.cpp:173-474) with everything it calls
 *   BamAlignment::TrimAlignment / TrimLowQualityEnds / TrimNumBases      src/bam_io.cpp:384-547
 *   AdapterTrimmer::trim_adapters / trim_five_prime / trim_three_prime   src/adapter_trimmer.cpp:54-176
 *   AlignmentFilters::GetEndDistToIndel / GetNumEndMatches / HasLargestEndMatches   src/alignment_filters.cpp
 *   BaseQuality::sum_log_prob_correct                                    src/base_quality.h:77-82
 *   BamProcessor::extract_mappings / get_valid_pairings                  src/bam_processor.cpp:59-157
 * and of remove_pcr_duplicates (src/pcr_duplicates.{h,cpp}).
 *
 * The Z-algorithm tables of the reference (src/zalgorithm.cpp) are replaced by direct prefix / suffix comparisons:
 * the windows are a few dozen positions wide, and the counts are defined by the strings alone.
 * Where the reference calls printErrorAndDie / assert, the functions here throw FilterError.
 */
#ifndef HIPSTR_B200_READ_FILTER_H_
#define HIPSTR_B200_READ_FILTER_H_

#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "bam_reader.h"

namespace hipstr {

struct FilterError : std::runtime_error {
  explicit FilterError(const std::string& what) : std::runtime_error(what) {}
};

/* BamAlignment::TrimAlignment (bam_io.cpp:384-477) */
void trim_alignment(BamRecord& a, int32_t min_read_start, int32_t max_read_stop, char min_base_qual = '~');
inline void trim_low_quality_ends(BamRecord& a, char min_base_qual) { trim_alignment(a, a.end_pos + 1, a.pos - 1, min_base_qual); }
void trim_num_bases(BamRecord& a, int left_trim, int right_trim);   /* :483-547 */

class AdapterTrimmer {
 public:
  AdapterTrimmer();   /* TruSeq + Nextera adapters, trimming on (adapter_trimmer.h:66-75) */
  void set_enabled(bool on) { trim_ = on; }
  void trim_adapters(BamRecord& a);
  int64_t trim_five_prime(BamRecord& a, const std::vector<std::string>& adapters) const;
  int64_t trim_three_prime(BamRecord& a, const std::vector<std::string>& adapters) const;
  int64_t r1_trimmed_bases = 0, r2_trimmed_bases = 0, r1_trimmed_reads = 0, r2_trimmed_reads = 0, r1_total_reads = 0, r2_total_reads = 0;
  std::string stats_message() const;   /* get_trimming_stats_msg */

 private:
  std::vector<std::string> r1_fw_, r2_fw_, r1_rc_, r2_rc_;
  bool trim_ = true;
};

namespace filters {
std::pair<int, int> end_dist_to_indel(const BamRecord& a);
std::pair<int, int> num_end_matches(const BamRecord& a, std::string_view ref_seq, int ref_seq_start);
bool has_largest_end_matches(const BamRecord& a, std::string_view ref_seq, int ref_seq_start, int max_external, int max_internal);
double sum_log_prob_correct(const std::string& quals);
}  // namespace filters

struct FilterOptions {          /* BamProcessor's public knobs with its defaults (bam_processor.h:78-101) */
  int32_t max_mate_dist = 1000;
  int32_t min_bp_before_indel = 7;
  int32_t min_flank = 5;
  int32_t min_read_end_match = 10;
  int32_t maximal_end_match_window = 15;
  int32_t require_paired_reads = 1;
  double min_sum_qual_log_prob = -10;
  int32_t max_total_reads = 1000000;
  char base_qual_trim = '5';
  bool remove_pcr_dups = true;
  bool trim_adapters = true;
};

struct FilterCounts {
  int32_t read_count = 0, hard_clip = 0, read_has_n = 0, low_qual_score = 0, unique_mapping = 0, unpaired_filtered = 0;
  int32_t pcr_duplicates = 0;
  bool too_many_reads = false;
};

struct FilteredReads {
  std::vector<std::string> rg_names;                                   /* sample of each read group, in order of appearance */
  std::vector<std::vector<BamRecord> > paired, mates, unpaired;        /* [read group] */
  FilterCounts counts;
};

class ReadFilter {
 public:
  FilterOptions options;
  AdapterTrimmer adapter_trimmer;
  /* records: the reader's stream for the padded region, files one after the other; ref_names / file_names resolve
   * ref_id / file; regions: [start, stop) of every STR of the group; rg_to_sample keys are file name + read group id. */
  void run(const std::vector<BamRecord>& records, const std::vector<std::string>& ref_names, const std::vector<std::string>& file_names,
           std::string_view chrom_seq, const std::vector<std::pair<int32_t, int32_t> >& regions,
           const std::map<std::string, std::string>& rg_to_sample, FilteredReads& out);
  /* remove_pcr_duplicates (pcr_duplicates.cpp:19-94); returns the number of duplicate sets removed */
  static int32_t remove_pcr_duplicates(const std::map<std::string, std::string>& rg_to_library, const std::vector<std::string>& file_names,
                                       FilteredReads& reads);

 private:
  void extract_mappings(const BamRecord& a, const std::vector<std::string>& ref_names, std::vector<std::pair<std::string, int32_t> >& out) const;
  void valid_pairings(const BamRecord& a1, const BamRecord& a2, const std::vector<std::string>& ref_names,
                      std::vector<std::pair<std::string, int32_t> >& p1, std::vector<std::pair<std::string, int32_t> >& p2) const;
};

std::string cigar_string(const std::vector<std::pair<char, int32_t> >& cigar);   /* BuildCigarString */

}  // namespace hipstr
#endif
