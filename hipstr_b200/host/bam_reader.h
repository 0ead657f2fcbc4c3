/*
 * bam_reader.h -- indexed BAM access for the reads of one region: host-side replacement of the reference's
 * BamCramReader / BamCramMultiReader / BamHeader / BamAlignment (src/bam_io.{h,cpp}) over htslib's BGZF, BAM and
 * BAI code (lib/htslib/{bgzf,sam,hts}.c, version 1.9) for the case the genotyper uses: BAM files with a .bai index,
 * read region by region, files visited one after the other (ORDER_ALNS_BY_FILE, bam_io.cpp:318-361).
 *
 * Written from the SAM/BAM specification (BGZF blocks, BAM records, the binning + linear index), not from htslib:
 * a region query decodes every record of the chunks the index names and keeps those that overlap
 * [start, end) on the chromosome, in file order -- the sequence hts_itr_next yields.  CRAM is out of scope.
 */
#ifndef HIPSTR_B200_BAM_READER_H_
#define HIPSTR_B200_BAM_READER_H_

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace hipstr {

/* What the reference keeps of an alignment (BamAlignment, bam_io.h:43-303), decoded eagerly. */
struct BamRecord {
  std::string name;
  uint16_t flag = 0;
  uint8_t mapq = 0;
  int32_t file = 0;                    /* index of the file it came from (Filename()) */
  int32_t ref_id = -1, mate_ref_id = -1;
  int32_t pos = 0, end_pos = -1;       /* Position(), GetEndPosition() (exclusive; bam_endpos) */
  int32_t mate_pos = -1;
  std::string bases, quals;            /* QueryBases(), Qualities() (Phred+33) */
  std::vector<std::pair<char, int32_t> > cigar;
  bool has_rg = false, has_xa = false, has_sa = false, has_as = false, has_xs = false, has_hp = false;
  std::string rg, xa, sa;              /* read group, alternate mappings (BWA XA), supplementary alignments (SA) */
  int64_t as = 0, xs = 0, hp = 0;      /* primary / suboptimal alignment score, 10X haplotype tag */
  std::string passes;                  /* the "PF" tag BamProcessor adds (bam_processor.cpp:21-27); empty = absent */
  int32_t length() const { return (int32_t)bases.size(); }
  bool paired() const { return flag & 0x1; }
  bool mapped() const { return !(flag & 0x4); }
  bool reverse() const { return flag & 0x10; }
  bool first_mate() const { return flag & 0x40; }
  bool second_mate() const { return flag & 0x80; }
};

struct BamReadGroup { std::string id, sample, library; bool has_sample = false, has_library = false; };

class BamFile {
 public:
  BamFile() {}
  ~BamFile();
  BamFile(const BamFile&) = delete;
  BamFile& operator=(const BamFile&) = delete;
  /* opens path and path + ".bai" (or path with .bam replaced by .bai); false + error() on failure */
  bool open(const std::string& path);
  const std::string& error() const { return error_; }
  const std::string& path() const { return path_; }
  const std::string& header_text() const { return text_; }
  const std::vector<std::string>& ref_names() const { return ref_names_; }
  const std::vector<uint32_t>& ref_lengths() const { return ref_lengths_; }
  const std::vector<BamReadGroup>& read_groups() const { return read_groups_; }   /* BamHeader::parse_read_groups */
  int ref_id(const std::string& name) const;
  /* appends the records overlapping [start, end) of chromosome `chrom`, in file order.  With `str_region` = {start, stop}
   * of the STR group, records that read_and_filter_reads drops on sight -- paired reads that are neither first nor second
   * mate, and reads that miss the STR whose mate cannot reach it either (src/bam_processor.cpp:191-203, decided by
   * position, length, flag and mate position alone) -- are skipped BEFORE their name, bases, qualities and tags are decoded. */
  bool fetch(const std::string& chrom, int32_t start, int32_t end, int32_t file_index, std::vector<BamRecord>& out,
             const int32_t* str_region = nullptr);

 private:
  struct Chunk { uint64_t beg, end; };
  struct RefIndex { std::map<uint32_t, std::vector<Chunk> > bins; std::vector<uint64_t> linear; };
  FILE* fp_ = nullptr;
  std::string path_, error_, text_;
  std::vector<std::string> ref_names_;
  std::vector<uint32_t> ref_lengths_;
  std::vector<BamReadGroup> read_groups_;
  std::vector<RefIndex> index_;
  /* one decompressed BGZF block */
  uint64_t block_addr_ = ~0ull;        /* file offset of the block in `block_` */
  uint64_t next_addr_ = 0;             /* file offset of the block after it */
  std::vector<unsigned char> block_;
  size_t block_at_ = 0;
  std::vector<unsigned char> record_;  /* the bytes of the record being decoded */

  bool fail(const std::string& why) { error_ = why; return false; }
  bool load_block(uint64_t addr);
  bool seek(uint64_t voffset);
  uint64_t tell() const { return (block_addr_ << 16) | (uint64_t)block_at_; }
  /* 1 = n bytes read, 0 = clean end of file before the first byte, -1 = error */
  int read(void* dst, size_t n);
  bool read_header();
  bool load_index(const std::string& path);
  int read_record(BamRecord& rec, const int32_t* str_region = nullptr);   /* 2 = skipped by the STR pre-filter */
};

}  // namespace hipstr
#endif
