/*
 * vcf_writer.h -- host-side mirror of the reference's VCFWriter (src/vcf_writer.h:31-84, src/vcf_writer.cpp:3-36):
 * same class and method names, same semantics.  Records of one chromosome may arrive up to MAX_RECORD_PAD = 50 bp
 * out of order (regions are processed in sorted order but a record's POS can precede its region's start by the
 * padding, vcf_writer.h:33-35); a min-heap keyed by POS holds them back until no later record can precede them.
 * Chromosomes must arrive grouped.  Output is BGZF (blocked gzip with the 'BC' extra field and the 28-byte EOF
 * block -- the format htslib's bgzf_write produces for the reference) when the path ends in ".gz", plain text
 * otherwise.  After a multi-GPU run rank 0 feeds the gathered records in (chromosome, position) order.
 */
#ifndef HIPSTR_B200_VCF_WRITER_H_
#define HIPSTR_B200_VCF_WRITER_H_

#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

namespace hipstr {

class VCFWriter {
 public:
  VCFWriter();
  ~VCFWriter();
  VCFWriter(const VCFWriter&) = delete;
  VCFWriter& operator=(const VCFWriter&) = delete;

  bool is_open() const { return open_; }
  bool open(const std::string& vcf_file);                 // false if the file cannot be created
  bool write_header(const std::string& header_text);      // false if not open
  bool add_vcf_record(const std::string& chrom, int32_t record_pos, const std::string& record_text);   // false once a write failed
  bool close();                                           // false if any write, the compression or fclose failed

 private:
  struct Record { int32_t pos; std::string text; };
  static bool later(const Record* a, const Record* b) { return a->pos > b->pos; }   // tuple_comparator
  void write_all_records();
  void emit(const std::string& s);       // appends to the current BGZF block / the plain file
  void flush_block();

  std::FILE* fp_;
  bool open_, bgzf_, io_ok_;
  std::string chrom_;
  std::vector<Record*> heap_;
  std::vector<unsigned char> block_;     // uncompressed bytes waiting for the next BGZF block
  int32_t max_record_pad_;
};

}  // namespace hipstr
#endif
