/* bam_reader.cpp -- see bam_reader.h. */
#include "bam_reader.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <sstream>

namespace hipstr {

namespace {

inline uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const unsigned char* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

/* the reference's HTSLIB_INT_TO_BASE (bam_io.h:22-25): only A, C, G, T, N survive, every other code is a blank */
const char kBase[16] = {' ', 'A', 'C', ' ', 'G', ' ', ' ', ' ', 'T', ' ', ' ', ' ', ' ', ' ', ' ', 'N'};
const char kCigarOp[] = "MIDNSHP=XB";

/* size in bytes of one auxiliary value of type `t` starting at p (end = first byte past the record), or -1 */
long aux_size(char t, const unsigned char* p, const unsigned char* end) {
  switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    case 'Z': case 'H': {
      const unsigned char* q = p;
      while (q < end && *q) q++;
      return q < end ? (long)(q - p) + 1 : -1;
    }
    case 'B': {
      if (end - p < 5) return -1;
      const long each = aux_size((char)p[0], p + 5, end);
      if (each < 1 || p[0] == 'Z' || p[0] == 'H' || p[0] == 'B') return -1;
      return 5 + each * (long)le32(p + 1);
    }
    default: return -1;
  }
}

bool aux_int(char t, const unsigned char* p, int64_t& v) {   /* bam_aux2i */
  switch (t) {
    case 'c': v = (int8_t)p[0]; return true;
    case 'C': v = p[0]; return true;
    case 's': v = (int16_t)le16(p); return true;
    case 'S': v = le16(p); return true;
    case 'i': v = (int32_t)le32(p); return true;
    case 'I': v = le32(p); return true;
    default: return false;
  }
}

/* bins that can hold records overlapping [beg, end): the UCSC binning scheme of the BAI index (SAM spec 5.3) */
void region_bins(int64_t beg, int64_t end, std::vector<uint32_t>& bins) {
  if (beg >= end) return;
  if (end > (1LL << 29)) end = 1LL << 29;
  --end;
  bins.push_back(0);
  const int shifts[5] = {26, 23, 20, 17, 14};
  const uint32_t first[5] = {1, 9, 73, 585, 4681};
  for (int l = 0; l < 5; l++)
    for (int64_t k = first[l] + (beg >> shifts[l]); k <= first[l] + (end >> shifts[l]); k++) bins.push_back((uint32_t)k);
}

}  // namespace

BamFile::~BamFile() {
  if (fp_) fclose(fp_);
}

bool BamFile::load_block(uint64_t addr) {
  if (addr == block_addr_) { block_at_ = 0; return true; }
  unsigned char head[18];
  if (fseeko(fp_, (off_t)addr, SEEK_SET) != 0) return fail("seek failed in " + path_);
  const size_t got = fread(head, 1, 18, fp_);
  if (got == 0) {   // end of file: an empty block
    block_addr_ = addr; next_addr_ = addr; block_.clear(); block_at_ = 0;
    return true;
  }
  if (got != 18 || head[0] != 31 || head[1] != 139 || head[2] != 8 || !(head[3] & 4)) return fail("not a BGZF block in " + path_);
  // the BC subfield holds the block size; it is the first (and in practice only) extra subfield
  const int xlen = le16(head + 10);
  std::vector<unsigned char> extra;
  int bsize = -1;
  if (xlen >= 6 && head[12] == 'B' && head[13] == 'C' && le16(head + 14) == 2) bsize = le16(head + 16);
  size_t consumed = 18;
  if (bsize < 0 || xlen != 6) {
    extra.resize((size_t)xlen);
    std::memcpy(extra.data(), head + 12, std::min<size_t>(6, (size_t)xlen));
    if (xlen > 6 && fread(extra.data() + 6, 1, (size_t)xlen - 6, fp_) != (size_t)xlen - 6) return fail("truncated BGZF block in " + path_);
    for (int at = 0; at + 4 <= xlen;) {
      const int len = le16(extra.data() + at + 2);
      if (extra[at] == 'B' && extra[at + 1] == 'C' && len == 2 && at + 6 <= xlen) bsize = le16(extra.data() + at + 4);
      at += 4 + len;
    }
    if (bsize < 0) return fail("BGZF block without a BC field in " + path_);
    consumed = 12 + (size_t)xlen;
  }
  const long payload = (long)bsize + 1 - (long)consumed - 8;
  if (payload < 0) return fail("bad BGZF block size in " + path_);
  std::vector<unsigned char> raw((size_t)payload + 8);
  if (fread(raw.data(), 1, raw.size(), fp_) != raw.size()) return fail("truncated BGZF block in " + path_);
  const uint32_t isize = le32(raw.data() + payload + 4);
  block_.resize(isize);
  if (isize) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return fail("zlib initialisation failed");
    zs.next_in = raw.data();
    zs.avail_in = (uInt)payload;
    zs.next_out = block_.data();
    zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != isize) return fail("corrupt BGZF block in " + path_);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), block_.data(), isize) != le32(raw.data() + payload)) return fail("BGZF checksum mismatch in " + path_);
  }
  block_addr_ = addr;
  next_addr_ = addr + (uint64_t)bsize + 1;
  block_at_ = 0;
  return true;
}

bool BamFile::seek(uint64_t voffset) {
  if (!load_block(voffset >> 16)) return false;
  block_at_ = (size_t)(voffset & 0xffff);
  return block_at_ <= block_.size() ? true : fail("virtual offset past its block in " + path_);
}

int BamFile::read(void* dst, size_t n) {
  unsigned char* out = (unsigned char*)dst;
  size_t done = 0;
  while (done < n) {
    if (block_at_ == block_.size()) {
      if (next_addr_ == block_addr_) return done == 0 ? 0 : -1;   // end of file
      if (!load_block(next_addr_)) return -1;
      continue;
    }
    const size_t take = std::min(n - done, block_.size() - block_at_);
    std::memcpy(out + done, block_.data() + block_at_, take);
    block_at_ += take;
    done += take;
  }
  return 1;
}

bool BamFile::read_header() {
  unsigned char b[8];
  if (read(b, 8) != 1 || std::memcmp(b, "BAM\1", 4) != 0) return fail(path_ + " is not a BAM file");
  const uint32_t l_text = le32(b + 4);
  text_.resize(l_text);
  if (l_text && read(&text_[0], l_text) != 1) return fail("truncated BAM header in " + path_);
  text_ = std::string(text_.c_str());   // the text may be NUL padded
  if (read(b, 4) != 1) return fail("truncated BAM header in " + path_);
  const uint32_t n_ref = le32(b);
  for (uint32_t i = 0; i < n_ref; i++) {
    if (read(b, 4) != 1) return fail("truncated BAM header in " + path_);
    std::string name(le32(b), '\0');
    if (name.empty() || read(&name[0], name.size()) != 1 || read(b, 4) != 1) return fail("truncated BAM header in " + path_);
    name.resize(name.size() - 1);
    ref_names_.push_back(name);
    ref_lengths_.push_back(le32(b));
  }
  // BamHeader::parse_read_groups (bam_io.cpp:44-65): every TAG:value field of the @RG lines
  std::stringstream ss(text_);
  std::string line;
  while (std::getline(ss, line)) {
    if (line.compare(0, 3, "@RG") != 0) continue;
    BamReadGroup rg;
    std::stringstream fields(line);
    std::string field;
    bool first = true;
    while (std::getline(fields, field, '\t')) {
      if (first) { first = false; continue; }
      const size_t colon = field.find(':');
      if (colon == std::string::npos) continue;
      const std::string tag = field.substr(0, colon), value = field.substr(colon + 1);
      if (tag == "ID") rg.id = value;
      else if (tag == "SM") { rg.sample = value; rg.has_sample = true; }
      else if (tag == "LB") { rg.library = value; rg.has_library = true; }
    }
    read_groups_.push_back(rg);
  }
  return true;
}

bool BamFile::load_index(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> data;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  fclose(f);
  size_t at = 0;
  auto need = [&](size_t k) { return at + k <= data.size(); };
  if (!need(8) || std::memcmp(data.data(), "BAI\1", 4) != 0) return fail(path + " is not a BAI index");
  const uint32_t n_ref = le32(data.data() + 4);
  at = 8;
  index_.assign(n_ref, RefIndex());
  for (uint32_t r = 0; r < n_ref; r++) {
    if (!need(4)) return fail("truncated index " + path);
    const uint32_t n_bin = le32(data.data() + at); at += 4;
    for (uint32_t b = 0; b < n_bin; b++) {
      if (!need(8)) return fail("truncated index " + path);
      const uint32_t bin = le32(data.data() + at), n_chunk = le32(data.data() + at + 4);
      at += 8;
      if (!need((size_t)n_chunk * 16)) return fail("truncated index " + path);
      if (bin != 37450) {   // 37450 is the metadata pseudo-bin
        std::vector<Chunk>& chunks = index_[r].bins[bin];
        for (uint32_t c = 0; c < n_chunk; c++) chunks.push_back(Chunk{le64(data.data() + at + 16 * c), le64(data.data() + at + 16 * c + 8)});
      }
      at += (size_t)n_chunk * 16;
    }
    if (!need(4)) return fail("truncated index " + path);
    const uint32_t n_intv = le32(data.data() + at); at += 4;
    if (!need((size_t)n_intv * 8)) return fail("truncated index " + path);
    for (uint32_t i = 0; i < n_intv; i++) index_[r].linear.push_back(le64(data.data() + at + 8 * i));
    at += (size_t)n_intv * 8;
  }
  return true;
}

bool BamFile::open(const std::string& path) {
  path_ = path;
  fp_ = fopen(path.c_str(), "rb");
  if (!fp_) return fail("File " + path + " does not exist");
  if (!read_header()) return false;
  if (!load_index(path + ".bai")) {
    if (!error_.empty()) return false;
    std::string alt = path;
    if (alt.size() > 4 && alt.compare(alt.size() - 4, 4, ".bam") == 0) alt.replace(alt.size() - 4, 4, ".bai");
    if (alt == path || !load_index(alt)) return error_.empty() ? fail("Failed to load the index for file " + path) : false;
  }
  if (index_.size() != ref_names_.size()) return fail("index of " + path + " does not match its header");
  return true;
}

int BamFile::ref_id(const std::string& name) const {
  for (size_t i = 0; i < ref_names_.size(); i++)
    if (ref_names_[i] == name) return (int)i;
  return -1;
}

int BamFile::read_record(BamRecord& rec, const int32_t* str_region) {
  unsigned char sz[4];
  const int rc = read(sz, 4);
  if (rc != 1) return rc;
  const uint32_t block_size = le32(sz);
  if (block_size < 32) { fail("invalid record in " + path_); return -1; }
  std::vector<unsigned char>& d = record_;   // reused from record to record
  if (d.size() < block_size) d.resize(block_size);
  if (read(d.data(), block_size) != 1) { fail("truncated record in " + path_); return -1; }
  const int32_t ref = (int32_t)le32(d.data()), pos = (int32_t)le32(d.data() + 4);
  const uint32_t l_name = d[8], n_cigar = le16(d.data() + 12), l_seq = le32(d.data() + 16);
  const size_t need = 32 + (size_t)l_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + l_seq;
  if (need > block_size || l_name == 0) { fail("invalid record in " + path_); return -1; }
  rec = BamRecord();
  rec.ref_id = ref;
  rec.pos = pos;
  rec.mapq = d[9];
  rec.flag = le16(d.data() + 14);
  rec.mate_ref_id = (int32_t)le32(d.data() + 20);
  rec.mate_pos = (int32_t)le32(d.data() + 24);
  const unsigned char* p = d.data() + 32;
  const unsigned char* name_at = p;
  p += l_name;
  int32_t ref_len = 0;
  for (uint32_t c = 0; c < n_cigar; c++, p += 4) {
    const uint32_t v = le32(p);
    const uint32_t op = v & 0xf;
    rec.cigar.emplace_back(op < 10 ? kCigarOp[op] : '?', (int32_t)(v >> 4));
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += (int32_t)(v >> 4);   // M D N = X consume the reference
  }
  rec.end_pos = (!(rec.flag & 0x4) && n_cigar > 0) ? pos + ref_len : pos + 1;   // bam_endpos
  if (l_seq == 0) rec.cigar.clear();   // BamAlignment::ExtractSequenceFields returns before the CIGAR when there is no sequence (bam_io.cpp:15-19)
  if (str_region) {   // the first two tests of read_and_filter_reads, which need none of the fields decoded below
    if (rec.paired() && !rec.first_mate() && !rec.second_mate()) return 2;
    if (rec.pos > str_region[1] || rec.end_pos < str_region[0]) {
      if (!rec.paired() || rec.mate_pos == rec.pos) return 2;
      if (rec.mate_pos > str_region[1]) return 2;
      if (rec.mate_pos + (int32_t)l_seq + 100 < str_region[0]) return 2;
    }
  }
  rec.name.assign((const char*)name_at, strnlen((const char*)name_at, l_name));
  rec.bases.resize(l_seq);
  for (uint32_t i = 0; i < l_seq; i++) rec.bases[i] = kBase[(p[i >> 1] >> ((~i & 1) << 2)) & 0xf];
  p += (l_seq + 1) / 2;
  rec.quals.resize(l_seq);
  for (uint32_t i = 0; i < l_seq; i++) rec.quals[i] = (char)(p[i] + 33);
  p += l_seq;
  const unsigned char* end = d.data() + block_size;
  while (end - p >= 3) {
    const char t0 = (char)p[0], t1 = (char)p[1], type = (char)p[2];
    p += 3;
    const long size = aux_size(type, p, end);
    if (size < 0 || p + size > end) { fail("invalid auxiliary field in " + path_); return -1; }
    const bool text = type == 'Z' || type == 'H';
    if (t0 == 'R' && t1 == 'G' && text) { rec.has_rg = true; rec.rg = (const char*)p; }
    else if (t0 == 'X' && t1 == 'A' && text) { rec.has_xa = true; rec.xa = (const char*)p; }
    else if (t0 == 'S' && t1 == 'A' && text) { rec.has_sa = true; rec.sa = (const char*)p; }
    else if (t0 == 'A' && t1 == 'S') rec.has_as = aux_int(type, p, rec.as);
    else if (t0 == 'X' && t1 == 'S') rec.has_xs = aux_int(type, p, rec.xs);
    else if (t0 == 'H' && t1 == 'P') rec.has_hp = aux_int(type, p, rec.hp);
    p += size;
  }
  return 1;
}

bool BamFile::fetch(const std::string& chrom, int32_t start, int32_t end, int32_t file_index, std::vector<BamRecord>& out,
                    const int32_t* str_region) {
  const int tid = ref_id(chrom);
  if (tid < 0) return fail("chromosome " + chrom + " is not in the header of " + path_);
  if (start < 0) start = 0;
  if (start >= end) return true;
  const RefIndex& idx = index_[tid];
  const size_t window = (size_t)start >> 14;
  const uint64_t min_off = window < idx.linear.size() ? idx.linear[window] : 0;
  std::vector<uint32_t> bins;
  region_bins(start, end, bins);
  std::vector<Chunk> chunks;
  for (uint32_t b : bins) {
    auto it = idx.bins.find(b);
    if (it == idx.bins.end()) continue;
    for (const Chunk& c : it->second)
      if (c.end > min_off) chunks.push_back(c);
  }
  std::sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.beg < b.beg; });
  std::vector<Chunk> merged;
  for (const Chunk& c : chunks) {
    if (!merged.empty() && c.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, c.end);
    else merged.push_back(c);
  }
  BamRecord rec;
  for (const Chunk& c : merged) {
    if (!seek(std::max(c.beg, min_off))) return false;
    while (tell() < c.end) {
      const int rc = read_record(rec, str_region);
      if (rc < 0) return false;
      if (rc == 0) break;
      if (rec.ref_id != tid || rec.pos >= end) return true;   // coordinate-sorted: nothing further can overlap
      if (rc == 1 && rec.end_pos > start) {
        rec.file = file_index;
        out.push_back(std::move(rec));
      }
    }
  }
  return true;
}

}  // namespace hipstr
