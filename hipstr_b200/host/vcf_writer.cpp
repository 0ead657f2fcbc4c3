#include "vcf_writer.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>

#include "../../include/hipstr_b200.h"

namespace hipstr {

namespace {
const size_t kBgzfBlock = 0xff00;   // uncompressed bytes per block, as htslib (BGZF_BLOCK_SIZE)
const unsigned char kBgzfEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                    0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
}

VCFWriter::VCFWriter() : fp_(nullptr), open_(false), bgzf_(false), io_ok_(true), max_record_pad_(50) {}
VCFWriter::~VCFWriter() { close(); }

bool VCFWriter::open(const std::string& vcf_file) {
  if (open_) return false;                       // the reference dies: "Cannot reopen an open VCFWriter"
  fp_ = std::fopen(vcf_file.c_str(), "wb");
  if (!fp_) return false;
  bgzf_ = vcf_file.size() >= 3 && vcf_file.compare(vcf_file.size() - 3, 3, ".gz") == 0;
  open_ = true;
  io_ok_ = true;
  chrom_.clear();
  return true;
}

void VCFWriter::flush_block() {
  if (block_.empty()) return;
  // one gzip member: header with the BC subfield, raw deflate payload, CRC32 + ISIZE
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { io_ok_ = false; block_.clear(); return; }
  std::vector<unsigned char> out(deflateBound(&zs, block_.size()) + 64);
  zs.next_in = block_.data();
  zs.avail_in = (uInt)block_.size();
  zs.next_out = out.data() + 18;
  zs.avail_out = (uInt)(out.size() - 18);
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) io_ok_ = false;
  const size_t clen = zs.total_out;
  deflateEnd(&zs);
  const unsigned char head[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0, 0};
  std::memcpy(out.data(), head, 18);
  const size_t total = 18 + clen + 8;            // BSIZE = total block size - 1
  out[16] = (unsigned char)((total - 1) & 0xff);
  out[17] = (unsigned char)((total - 1) >> 8);
  const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), block_.data(), (uInt)block_.size());
  const uint32_t isize = (uint32_t)block_.size();
  for (int k = 0; k < 4; k++) {
    out[18 + clen + k] = (unsigned char)(crc >> (8 * k));
    out[18 + clen + 4 + k] = (unsigned char)(isize >> (8 * k));
  }
  if (std::fwrite(out.data(), 1, total, fp_) != total) io_ok_ = false;   // full disk, closed pipe, ...
  block_.clear();
}

void VCFWriter::emit(const std::string& s) {
  if (!bgzf_) { if (std::fwrite(s.data(), 1, s.size(), fp_) != s.size()) io_ok_ = false; return; }
  size_t at = 0;
  while (at < s.size()) {
    const size_t take = std::min(s.size() - at, kBgzfBlock - block_.size());
    block_.insert(block_.end(), s.begin() + at, s.begin() + at + take);
    at += take;
    if (block_.size() == kBgzfBlock) flush_block();
  }
}

bool VCFWriter::write_header(const std::string& header_text) {
  if (!open_) return false;
  emit(header_text);
  return io_ok_;
}

void VCFWriter::write_all_records() {
  while (!heap_.empty()) {
    std::pop_heap(heap_.begin(), heap_.end(), later);
    Record* best = heap_.back();
    heap_.pop_back();
    emit(best->text + "\n");
    delete best;
  }
}

bool VCFWriter::add_vcf_record(const std::string& chrom, int32_t record_pos, const std::string& record_text) {
  if (!open_) return false;
  if (chrom != chrom_) {            // new chromosome: everything held back goes out first
    write_all_records();
    chrom_ = chrom;
  } else {
    while (!heap_.empty()) {        // records that precede every possible future record
      std::pop_heap(heap_.begin(), heap_.end(), later);
      Record* best = heap_.back();
      heap_.pop_back();
      if (best->pos < record_pos - max_record_pad_) {
        emit(best->text + "\n");
        delete best;
      } else {
        heap_.push_back(best);
        std::push_heap(heap_.begin(), heap_.end(), later);
        break;
      }
    }
  }
  heap_.push_back(new Record{record_pos, record_text});
  std::push_heap(heap_.begin(), heap_.end(), later);
  return io_ok_;   // false once any write failed: the file on disk is incomplete
}

bool VCFWriter::close() {
  if (!open_) return io_ok_;
  write_all_records();
  if (bgzf_) {
    flush_block();
    if (std::fwrite(kBgzfEof, 1, sizeof(kBgzfEof), fp_) != sizeof(kBgzfEof)) io_ok_ = false;
  }
  if (std::fclose(fp_) != 0) io_ok_ = false;   // buffered data that could not be written shows up here
  fp_ = nullptr;
  open_ = false;
  return io_ok_;
}

}  // namespace hipstr

// C-ABI wrappers (include/hipstr_b200.h)
extern "C" {
hipstr_vcf_writer_t* hipstr_vcf_writer_open(const char* path) {
  if (!path) return nullptr;
  hipstr::VCFWriter* w = new hipstr::VCFWriter();
  if (!w->open(path)) { delete w; return nullptr; }
  return reinterpret_cast<hipstr_vcf_writer_t*>(w);
}
hipstr_status_t hipstr_vcf_writer_header(hipstr_vcf_writer_t* w, const char* text) {
  if (!w || !text) return HIPSTR_ERR_BAD_ARG;
  return reinterpret_cast<hipstr::VCFWriter*>(w)->write_header(text) ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_vcf_writer_add_record(hipstr_vcf_writer_t* w, const char* chrom, int32_t pos, const char* text) {
  if (!w || !chrom || !text) return HIPSTR_ERR_BAD_ARG;
  return reinterpret_cast<hipstr::VCFWriter*>(w)->add_vcf_record(chrom, pos, text) ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
void hipstr_vcf_writer_close(hipstr_vcf_writer_t* w) { hipstr_vcf_writer_finish(w); }
hipstr_status_t hipstr_vcf_writer_finish(hipstr_vcf_writer_t* w) {
  if (!w) return HIPSTR_ERR_BAD_ARG;
  hipstr::VCFWriter* p = reinterpret_cast<hipstr::VCFWriter*>(w);
  const bool ok = p->close();
  delete p;
  return ok ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
}
