/*
 * ref_bam_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" adapter around the UNMODIFIED reference code UPSTREAM of the genotyper (SURVEY.md section 8(f)
 * row 4): BAM access (src/bam_io.cpp over the vendored htslib 1.9, compiled file by file by oracle/Makefile),
 * BamProcessor::read_and_filter_reads (src/bam_processor.cpp:173-474) with its AdapterTrimmer /
 * AlignmentFilters / pairing logic, remove_pcr_duplicates (src/pcr_duplicates.cpp) and calc_het_snp_factors
 * (src/snp_phasing_quality.cpp) over a real SNPTree (src/snp_tree.h).  `#define private public` only lets the
 * harness call private members and fill a BamAlignment in memory; no reference source is modified or copied.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "bam_io.h"
#include "bam_processor.h"
#include "genotyper_bam_processor.h"
#undef private
#undef protected
#include "alignment_filters.h"
#include "mathops.h"
#include "base_quality.h"
#include "pcr_duplicates.h"
#include "region.h"
#include "snp_phasing_quality.h"
#include "snp_tree.h"
#include "vcf_reader.h"
#include "vcf_input.h"
extern "C" {
#include "htslib/htslib/tbx.h"
}

namespace {

void fill(BamAlignment& b, int32_t pos, int32_t end_pos, const std::string& bases, const std::string& quals, int32_t n_cigar,
          const char* cigar_type, const int32_t* cigar_len) {
  b.bases_ = bases;
  b.qualities_ = quals;
  b.cigar_ops_.clear();
  for (int i = 0; i < n_cigar; i++) b.cigar_ops_.push_back(CigarOp(cigar_type[i], cigar_len[i]));
  b.built_ = true;
  b.length_ = (int32_t)bases.size();
  b.pos_ = pos;
  b.end_pos_ = end_pos;
}

/* the two pure virtuals of BamProcessor that read_and_filter_reads never reaches */
class HarnessProcessor : public BamProcessor {
 public:
  HarnessProcessor(bool use_bam_rgs, bool remove_pcr_dups) : BamProcessor(use_bam_rgs, remove_pcr_dups) {}
  void verify_vcf_chromosomes(const std::vector<std::string>&) {}
  void init_output_vcf(const std::string&, const std::vector<std::string>&, const std::string&) {}
  void process_reads(std::vector<BamAlnList>&, std::vector<BamAlnList>&, std::vector<BamAlnList>&, const std::vector<std::string>&,
                     const RegionGroup&, const std::string&) {}
};

void dump(std::ostringstream& out, const char* kind, BamAlignment& a) {
  std::string pf;
  if (!a.GetStringTag("PF", pf)) pf = "-";
  out << kind << '\t' << a.Name() << '\t' << a.b_->core.flag << '\t' << a.Position() << '\t' << a.GetEndPosition() << '\t'
      << (a.CigarData().empty() ? std::string("*") : BuildCigarString(a.CigarData())) << '\t' << a.QueryBases() << '\t'
      << a.Qualities() << '\t' << pf << '\n';
}

}  // namespace

extern "C" {

/* calc_het_snp_factors for every entry (same flat layout as hipstr_snp_phasing_t); totals = {match_count, mismatch_count}. */
void ref_snp_phasing(int32_t n_entries, const int32_t* entry_aln_off, const int32_t* entry_snp_set, const int32_t* aln_pos,
                     const int32_t* aln_end, const int32_t* aln_seq_off, const char* bases, const char* quals,
                     const int32_t* aln_cigar_off, const char* cigar_type, const int32_t* cigar_len, int32_t n_sets,
                     const int32_t* set_off, const uint32_t* snp_pos, const char* base1, const char* base2, double* log_p1,
                     double* log_p2, int32_t* entry_counts /* [n_entries][2] */) {
  BaseQuality base_quality;
  std::vector<SNPTree*> trees;
  for (int s = 0; s < n_sets; s++) {
    std::vector<SNP> snps;
    for (int i = set_off[s]; i < set_off[s + 1]; i++) snps.push_back(SNP(snp_pos[i], base1[i], base2[i]));
    trees.push_back(new SNPTree(snps));
  }
  for (int e = 0; e < n_entries; e++) {
    log_p1[e] = log_p2[e] = 0;
    entry_counts[2 * e] = entry_counts[2 * e + 1] = 0;
    if (entry_snp_set[e] < 0) continue;
    std::vector<BamAlignment> alns(entry_aln_off[e + 1] - entry_aln_off[e]);
    for (size_t k = 0; k < alns.size(); k++) {
      const int a = entry_aln_off[e] + (int)k;
      fill(alns[k], aln_pos[a], aln_end[a], std::string(bases + aln_seq_off[a], bases + aln_seq_off[a + 1]),
           std::string(quals + aln_seq_off[a], quals + aln_seq_off[a + 1]), aln_cigar_off[a + 1] - aln_cigar_off[a],
           cigar_type + aln_cigar_off[a], cigar_len + aln_cigar_off[a]);
    }
    std::vector<double> p1, p2;
    int32_t match = 0, mismatch = 0;
    if (alns.size() == 2) {
      std::vector<BamAlignment> str_reads(1, alns[0]), mates(1, alns[1]);
      calc_het_snp_factors(str_reads, mates, base_quality, trees[entry_snp_set[e]], p1, p2, match, mismatch);
    } else {
      std::vector<BamAlignment> str_reads(1, alns[0]);
      calc_het_snp_factors(str_reads, base_quality, trees[entry_snp_set[e]], p1, p2, match, mismatch);
    }
    log_p1[e] = p1[0];
    log_p2[e] = p2[0];
    entry_counts[2 * e] = match;
    entry_counts[2 * e + 1] = mismatch;
  }
  for (SNPTree* t : trees) delete t;
}

/* SAM text file -> BAM + .bai with htslib, so that tests can hand the reference real files. Returns 0 on success. */
int32_t ref_sam_to_bam(const char* sam_path, const char* bam_path) {
  samFile* in = sam_open(sam_path, "r");
  if (!in) return -1;
  bam_hdr_t* hdr = sam_hdr_read(in);
  if (!hdr) { sam_close(in); return -2; }
  samFile* out = sam_open(bam_path, "wb");
  if (!out) { bam_hdr_destroy(hdr); sam_close(in); return -3; }
  int rc = 0;
  if (sam_hdr_write(out, hdr) < 0) rc = -4;
  bam1_t* b = bam_init1();
  int r;
  while (rc == 0 && (r = sam_read1(in, hdr, b)) >= 0)
    if (sam_write1(out, hdr, b) < 0) rc = -5;
  if (rc == 0 && r < -1) rc = -6;
  bam_destroy1(b);
  bam_hdr_destroy(hdr);
  sam_close(in);
  if (sam_close(out) < 0 && rc == 0) rc = -7;
  if (rc == 0 && sam_index_build(bam_path, 0) < 0) rc = -8;
  return rc;
}

/* All alignments BamCramMultiReader yields for SetRegion(chrom, start, end), one text line each
 * (name, flag, tid-name, pos, end, mapq, cigar, mate ref, mate pos, bases, quals, RG, XA, SA, AS, XS with "-" when
 * absent): the checker of the product's BAM decoding + index query. */
int32_t ref_bam_region_reads(int32_t n_files, const char* const* paths, const char* chrom, int32_t start, int32_t end,
                             int32_t cap, char* out_text) {
  std::vector<std::string> files(paths, paths + n_files);
  BamCramMultiReader reader(files, "", BamCramMultiReader::ORDER_ALNS_BY_FILE);
  std::ostringstream out;
  if (!reader.SetRegion(chrom, start, end)) return -1;
  BamAlignment a;
  while (reader.GetNextAlignment(a)) {
    std::string rg = "-", xa = "-", sa = "-";
    a.GetStringTag("RG", rg); a.GetStringTag("XA", xa); a.GetStringTag("SA", sa);
    int64_t as = 0, xs = 0;
    const bool has_as = a.HasTag("AS") && a.GetIntTag("AS", as), has_xs = a.HasTag("XS") && a.GetIntTag("XS", xs);
    out << a.Name() << '\t' << a.b_->core.flag << '\t' << a.Ref() << '\t' << a.Position() << '\t' << a.GetEndPosition() << '\t'
        << a.MapQuality() << '\t' << (a.CigarData().empty() ? std::string("*") : BuildCigarString(a.CigarData())) << '\t'
        << a.MateRef() << '\t' << a.MatePosition() << '\t' << a.QueryBases() << '\t' << a.Qualities() << '\t' << rg << '\t' << xa
        << '\t' << sa << '\t';
    if (has_as) out << as; else out << '-';
    out << '\t';
    if (has_xs) out << xs; else out << '-';
    out << '\t' << a.Filename() << '\n';
  }
  const std::string s = out.str();
  if ((int32_t)s.size() + 1 > cap) return -2;
  memcpy(out_text, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}

/* BamProcessor::process_regions for ONE region up to (not including) process_reads (bam_processor.cpp:551-607):
 * SetRegion with the MAX_MATE_DIST padding, read_and_filter_reads, optionally remove_pcr_duplicates.  Read groups map to
 * samples / libraries through rg_keys[i] = file name + read group id (what main builds from the BAM headers).
 * options = {MIN_FLANK, MIN_READ_END_MATCH, MAXIMAL_END_MATCH_WINDOW, MIN_BP_BEFORE_INDEL, REQUIRE_PAIRED_READS,
 *            BASE_QUAL_TRIM (char code), MAX_TOTAL_READS, MAX_MATE_DIST, remove PCR duplicates, trim adapters};
 * min_sum_qual_log_prob is MIN_SUM_QUAL_LOG_PROB.  Output text: a "G <sample>" line per read group in rg_names order, followed by
 * its lists, one line per alignment ("P" STR read of a pair, "M" its mate, "U" unpaired); the last line is
 * "T <TOO_MANY_READS>". */
int32_t ref_read_and_filter(int32_t n_files, const char* const* paths, const char* chrom, const char* chrom_seq, int32_t region_start,
                            int32_t region_stop, int32_t period, int32_t n_rg, const char* const* rg_keys,
                            const char* const* rg_samples, const char* const* rg_libraries, const int32_t* options,
                            double min_sum_qual_log_prob, int32_t cap, char* out_text) {
  std::vector<std::string> files(paths, paths + n_files);
  BamCramMultiReader reader(files, "", BamCramMultiReader::ORDER_ALNS_BY_FILE);
  HarnessProcessor proc(true, options[8] != 0);
  proc.suppress_all_logging();
  proc.MIN_FLANK = options[0];
  proc.MIN_READ_END_MATCH = options[1];
  proc.MAXIMAL_END_MATCH_WINDOW = options[2];
  proc.MIN_BP_BEFORE_INDEL = options[3];
  proc.REQUIRE_PAIRED_READS = options[4];
  proc.BASE_QUAL_TRIM = (char)options[5];
  proc.MAX_TOTAL_READS = options[6];
  proc.MAX_MATE_DIST = options[7];
  proc.MIN_SUM_QUAL_LOG_PROB = min_sum_qual_log_prob;
  if (!options[9]) proc.adapter_trimmer_.trim_ = false;
  std::map<std::string, std::string> rg_to_sample, rg_to_library;
  for (int i = 0; i < n_rg; i++) { rg_to_sample[rg_keys[i]] = rg_samples[i]; rg_to_library[rg_keys[i]] = rg_libraries[i]; }
  const std::string seq(chrom_seq);
  Region region(chrom, region_start, region_stop, period);
  RegionGroup group(region);
  if (!reader.SetRegion(chrom, region_start < proc.MAX_MATE_DIST ? 0 : region_start - proc.MAX_MATE_DIST, region_stop + proc.MAX_MATE_DIST))
    return -1;
  std::vector<std::string> rg_names;
  std::vector<std::vector<BamAlignment> > paired, mates, unpaired;
  proc.read_and_filter_reads(reader, seq, group, rg_to_sample, rg_names, paired, mates, unpaired, NULL, NULL);
  if (proc.REMOVE_PCR_DUPS == 1) {
    std::ostringstream sink;
    remove_pcr_duplicates(proc.base_quality_, true, rg_to_library, paired, mates, unpaired, sink);
  }
  std::ostringstream out;
  for (size_t g = 0; g < rg_names.size(); g++) {
    out << "G\t" << rg_names[g] << '\n';
    for (size_t i = 0; i < paired[g].size(); i++) { dump(out, "P", paired[g][i]); dump(out, "M", mates[g][i]); }
    for (size_t i = 0; i < unpaired[g].size(); i++) dump(out, "U", unpaired[g][i]);
  }
  out << "T\t" << (proc.TOO_MANY_READS ? 1 : 0) << '\n';
  const std::string s = out.str();
  if ((int32_t)s.size() + 1 > cap) return -2;
  memcpy(out_text, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}

/* The single-read filters and trimmers, on an alignment assembled in memory (flag only matters for adapter trimming):
 * what = 0 TrimLowQualityEnds(arg), 1 AdapterTrimmer::trim_adapters (default adapters), 2 TrimNumBases(arg, arg2).
 * The alignment after the step is returned like ref_left_align_one does. */
int32_t ref_trim_one(int32_t what, int32_t arg, int32_t arg2, int32_t flag, int32_t pos, int32_t end_pos, const char* bases,
                     const char* quals, int32_t n_cigar, const char* cigar_type, const int32_t* cigar_len, int32_t* out_pos,
                     char* out_seq, char* out_qual, int32_t* n_out_cigar, char* out_ctype, int32_t* out_clen) {
  BamAlignment b;
  fill(b, pos, end_pos, bases, quals, n_cigar, cigar_type, cigar_len);
  b.b_->core.flag = (uint16_t)flag;
  int32_t rc = 0;
  if (what == 0) b.TrimLowQualityEnds((char)arg);
  else if (what == 1) { AdapterTrimmer trimmer; trimmer.trim_adapters(b); }
  else b.TrimNumBases(arg, arg2);
  out_pos[0] = b.Position();
  out_pos[1] = b.GetEndPosition();
  out_pos[2] = b.Length();
  strcpy(out_seq, b.QueryBases().c_str());
  strcpy(out_qual, b.Qualities().c_str());
  *n_out_cigar = (int32_t)b.CigarData().size();
  for (int i = 0; i < *n_out_cigar; i++) { out_ctype[i] = b.CigarData()[i].Type; out_clen[i] = b.CigarData()[i].Length; }
  return rc;
}

/* AlignmentFilters on an alignment assembled in memory: out = {HasLargestEndMatches(aln, ref, 0, window, window),
 * GetNumEndMatches.first, .second, GetEndDistToIndel.first, .second}; sum_qual = BaseQuality::sum_log_prob_correct. */
void ref_alignment_filters(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar, const char* cigar_type,
                           const int32_t* cigar_len, const char* chrom_seq, int32_t window, int32_t* out, double* sum_qual) {
  BamAlignment b;
  fill(b, pos, end_pos, bases, quals, n_cigar, cigar_type, cigar_len);
  const std::string seq(chrom_seq);
  out[0] = AlignmentFilters::HasLargestEndMatches(b, seq, 0, window, window) ? 1 : 0;
  std::pair<int, int> m = AlignmentFilters::GetNumEndMatches(b, seq, 0);
  out[1] = m.first; out[2] = m.second;
  std::pair<int, int> d = AlignmentFilters::GetEndDistToIndel(b);
  out[3] = d.first; out[4] = d.second;
  BaseQuality base_quality;
  *sum_qual = base_quality.sum_log_prob_correct(b.Qualities());
}

/* The reference program from its BAM files to its VCF, as hipstr_main.cpp:360-555 runs it (minus option parsing): a
 * GenotyperBamProcessor over a BamCramMultiReader, read groups taken from the BAM headers (library tag "LB"),
 * process_regions() over the region file, finish().  No SNP VCF, no reference-panel VCF.
 * options = {use the default stutter model 0.95/0.05/0.05/0.95/0.01/0.01 instead of EM training, MIN_TOTAL_READS,
 *            REMOVE_PCR_DUPS, REQUIRE_PAIRED_READS, recalc_stutter_model_ (0/1), output GLs, output PLs, output FILTERS,
 *            treat chr1 as haploid (--haploid-chrs chr1), BAMs carry 10X haplotype tags (--10x-bams)}. */
int32_t ref_process_regions(int32_t n_files, const char* const* paths, const char* fasta_path, const char* region_path,
                            const char* vcf_out_path, const int32_t* options, const char* snp_vcf_path /* NULL = none */,
                            const char* ref_vcf_path /* NULL = none */) {
  std::vector<std::string> files(paths, paths + n_files);
  precompute_integer_logs();   // hipstr_main.cpp:352
  GenotyperBamProcessor proc(true, options[2] != 0);
  proc.suppress_all_logging();
  proc.MIN_TOTAL_READS = options[1];
  proc.REQUIRE_PAIRED_READS = options[3];
  if (options[0]) proc.set_default_stutter_model(0.95, 0.05, 0.05, 0.95, 0.01, 0.01);
  proc.recalc_stutter_model_ = options[4] != 0;
  Genotyper::OUTPUT_GLS = options[5];
  Genotyper::OUTPUT_PLS = options[6];
  Genotyper::OUTPUT_FILTERS = options[7];
  if (options[8]) proc.add_haploid_chrom("chr1");
  if (options[9]) proc.use_10x_bam_tags();
  BamCramMultiReader reader(files, "", BamCramMultiReader::ORDER_ALNS_BY_FILE);
  std::map<std::string, std::string> rg_to_sample, rg_to_library;
  std::set<std::string> samples;
  for (size_t i = 0; i < files.size(); i++) {
    const std::vector<ReadGroup>& groups = reader.bam_header()->read_groups(i);
    for (auto rg = groups.begin(); rg != groups.end(); ++rg) {
      rg_to_sample[files[i] + rg->GetID()] = rg->GetSample();
      rg_to_library[files[i] + rg->GetID()] = rg->GetLibrary();
      samples.insert(rg->GetSample());
    }
  }
  if (snp_vcf_path != NULL) proc.set_input_snp_vcf(snp_vcf_path);
  if (ref_vcf_path != NULL) proc.set_ref_vcf(ref_vcf_path);
  proc.set_output_str_vcf(vcf_out_path, fasta_path, "harness", samples);
  proc.process_regions(reader, region_path, fasta_path, rg_to_sample, rg_to_library, "harness", NULL, NULL, 10000000, "");
  proc.finish();
  return 0;
}

/* Plain VCF text -> bgzipped VCF + tabix index with htslib (so that VCF::VCFReader can open it). */
int32_t ref_vcf_bgzip_tabix(const char* vcf_text_path, const char* out_gz_path) {
  FILE* in = fopen(vcf_text_path, "rb");
  if (!in) return -1;
  BGZF* out = bgzf_open(out_gz_path, "w");
  if (!out) { fclose(in); return -2; }
  char buf[1 << 16];
  size_t n;
  int rc = 0;
  while ((n = fread(buf, 1, sizeof(buf), in)) > 0)
    if (bgzf_write(out, buf, n) < 0) rc = -3;
  fclose(in);
  if (bgzf_close(out) < 0 && rc == 0) rc = -4;
  if (rc == 0 && tbx_index_build(out_gz_path, 0, &tbx_conf_vcf) != 0) rc = -5;
  return rc;
}

/* create_snp_trees (src/snp_tree.cpp:26-108, no pedigree) for one region as SNPBamProcessor::process_reads calls it
 * (src/snp_bam_processor.cpp:62-63), dumped per VCF sample by querying each tree for everything: lines "S <sample>"
 * followed by "pos base1 base2".  Returns the text length, -1 when create_snp_trees fails (chromosome not in the VCF). */
int32_t ref_snp_sets(const char* snp_vcf_path, const char* chrom, int32_t region_start, int32_t region_stop, int32_t period,
                     int32_t max_mate_dist, int32_t skip_padding, int32_t cap, char* out_text) {
  VCF::VCFReader reader(snp_vcf_path);
  Region region(chrom, region_start, region_stop, period);
  std::vector<Region> skip(1, region);
  std::vector<SNPTree*> trees;
  std::map<std::string, unsigned int> sample_indices;
  std::ostringstream sink, out;
  if (!create_snp_trees(chrom, (region_start > max_mate_dist ? region_start - max_mate_dist : 1), region_stop + max_mate_dist, skip,
                        skip_padding, &reader, NULL, sample_indices, trees, sink))
    return -1;
  const std::vector<std::string>& names = reader.get_samples();
  for (size_t i = 0; i < names.size(); i++) {
    out << "S " << names[i] << '\n';
    std::vector<SNP> snps;
    trees[sample_indices[names[i]]]->findContained(0, 2000000000u, snps);
    for (size_t k = 0; k < snps.size(); k++) out << snps[k].pos() << ' ' << snps[k].base_one() << ' ' << snps[k].base_two() << '\n';
  }
  destroy_snp_trees(trees);
  const std::string s = out.str();
  if ((int32_t)s.size() + 1 > cap) return -2;
  memcpy(out_text, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}

/* read_vcf_alleles (src/vcf_input.cpp:21-50) on a bgzipped + tabix-indexed panel: returns 1 and pos / the alleles (one per line), or 0. */
int32_t ref_read_vcf_alleles(const char* vcf_path, const char* chrom, int32_t region_start, int32_t region_stop, int32_t period, int32_t* pos,
                             int32_t cap, char* out_text) {
  VCF::VCFReader reader(vcf_path);
  Region region(chrom, region_start, region_stop, period);
  std::vector<std::string> alleles;
  if (!read_vcf_alleles(&reader, region, alleles, *pos)) return 0;
  std::string s;
  for (size_t i = 0; i < alleles.size(); i++) { s += alleles[i]; s += '\n'; }
  if ((int32_t)s.size() + 1 > cap) return -2;
  memcpy(out_text, s.c_str(), s.size() + 1);
  return 1;
}

}  // extern "C"
