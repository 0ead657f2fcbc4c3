/*
 * haplotype_generator.h -- candidate haplotype blocks of one locus from its left-aligned reads: host-side mirror of the
 * reference's HaplotypeGenerator (src/SeqAlignment/HaplotypeGenerator.{h,cpp}) as driven by
 * SeqStutterGenotyper::build_haplotype (src/seq_stutter_genotyper.cpp:422-484) when no reference VCF is given.
 * Per STR region: the region padded by 5 bp is cut out of every spanning read, sequences with enough read / sample
 * support become the alleles of a repeat block (reference first, the others by length then sequence), identical
 * padding is trimmed again, and the block is fused with reference-only flank blocks of at most 35 bp.
 */
#ifndef HIPSTR_B200_HAPLOTYPE_GENERATOR_H_
#define HIPSTR_B200_HAPLOTYPE_GENERATOR_H_

#include <stdint.h>

#include <string>
#include <string_view>
#include <vector>

#include "seq_stutter_genotyper.h"

namespace hipstr {

/* A left-aligned read as the generator sees it (Alignment, SeqAlignment/AlignmentData.h:29-137). */
struct ReadView {
  int32_t start, stop;            /* Alignment::get_start() / get_stop() (HipSTR: stop = last aligned reference position) */
  const char* bases;              /* ungapped read sequence */
  int32_t n_cigar;
  const char* cigar_type;         /* '=', 'X', 'I', 'D' */
  const int32_t* cigar_len;
};

class HaplotypeGenerator {
 public:
  HaplotypeGenerator(int32_t min_aln_start, int32_t max_aln_stop) : min_aln_start_(min_aln_start), max_aln_stop_(max_aln_stop) {}
  /* add_haplotype_block (HaplotypeGenerator.cpp:274-329); reads grouped by sample */
  bool add_haplotype_block(int32_t region_start, int32_t region_stop, int32_t period, std::string_view chrom_seq,
                           const std::vector<std::vector<ReadView> >& alignments, const double* stutter);
  /* add_vcf_haplotype_block (:256-284): the alleles come from a reference panel instead of the reads */
  bool add_vcf_haplotype_block(int32_t pos, int32_t period, std::string_view chrom_seq, const std::vector<std::string>& vcf_alleles,
                               const double* stutter);
  bool fuse_haplotype_blocks(std::string_view chrom_seq);   /* :331-366 */
  const std::string& failure_msg() const { return failure_msg_; }
  const std::vector<HapBlock>& get_haplotype_blocks() const { return hap_blocks_; }

  /* the sequence of [region_start, region_end) in a read that spans it (extract_sequence, :82-156) */
  static bool extract_sequence(const ReadView& aln, int32_t region_start, int32_t region_end, std::string& seq);

 private:
  int32_t min_aln_start_, max_aln_stop_;
  std::string failure_msg_;
  std::vector<HapBlock> hap_blocks_;
  void gen_candidate_seqs(const std::string& ref_seq, int ideal_min_length, const std::vector<std::vector<ReadView> >& alignments,
                          int32_t& region_start, int32_t& region_end, std::vector<std::string>& sequences) const;
  void trim(int ideal_min_length, int32_t& region_start, int32_t& region_end, std::vector<std::string>& sequences) const;
};

}  // namespace hipstr
#endif
