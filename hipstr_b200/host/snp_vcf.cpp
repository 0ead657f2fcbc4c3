/*
 * snp_vcf.cpp -- the phased SNP VCF behind the phasing log-likelihoods: host-side replacement of VCF::VCFReader +
 * VCF::Variant (src/vcf_reader.{h,cpp}, over htslib's tabix / VCF code) as create_snp_trees uses them
 * (src/snp_tree.cpp:26-108, called from SNPBamProcessor::process_reads, src/snp_bam_processor.cpp:60-64).
 *
 * Different design: the file (bgzipped or plain) is read ONCE; its biallelic SNPs and every sample's phased heterozygous
 * calls are kept per chromosome, sorted by position, so the SNPs of a region are a binary search instead of a tabix
 * query + text parse per region.  What a region query returns is what create_snp_trees puts into its per-sample SNPTrees:
 *   records with start <= POS <= end (tabix region "chrom:start-end"), n_allele == 2 and bcf_is_snp (every allele one
 *   character), not within `padding` of a region to skip (in_any_region, snp_tree.cpp:9-15); per sample the calls that are
 *   not missing, phased ('|' before the second allele) and heterozygous, as SNP(POS - 1, allele[gt_a][0], allele[gt_b][0]).
 * Pedigree-based filtering (HaplotypeTracker) is not built.  Samples whose GT is not diploid are treated as missing.
 */
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hipstr_b200.h"

namespace {

struct Site {
  int32_t pos;        // 1-based POS
  char ref, alt;
  size_t calls;       // offset of this site's per-sample codes (sites x samples exceeds 32 bits on large panels)
};
struct Chrom {
  std::vector<Site> sites;          // in file order; sorted by position on first use
  std::vector<uint8_t> codes;       // [site][sample]: 0 unusable, 1 = 0|1 (ref on haplotype one), 2 = 1|0
  bool sorted = false;
};

/* Streams a plain or BGZF / gzip text file line by line: the file is read and inflated in bounded chunks (1 MiB of
 * compressed input, 4 MiB of text at a time), so memory stays proportional to the PARSED data, a file of any size works
 * (zlib's 32-bit avail_in is fed in pieces), and a stream that ends inside a gzip member is an error instead of a
 * silently shorter file. */
class LineStream {
 public:
  LineStream(const std::string& path, std::string& err) : err_(err), in_(1 << 20), out_(4 << 20) {
    f_ = fopen(path.c_str(), "rb");
    if (!f_) { err_ = "Failed to open the VCF file " + path; failed_ = true; return; }
    path_ = path;
    const size_t n = fread(in_.data(), 1, in_.size(), f_);
    in_len_ = n;
    gz_ = n >= 2 && (unsigned char)in_[0] == 31 && (unsigned char)in_[1] == 139;
    if (gz_) {
      std::memset(&zs_, 0, sizeof(zs_));
      if (inflateInit2(&zs_, 15 + 16) != Z_OK) { err_ = "zlib initialisation failed"; failed_ = true; return; }
      zs_.next_in = (Bytef*)in_.data();
      zs_.avail_in = (uInt)in_len_;
      z_open_ = true;
    }
  }
  ~LineStream() {
    if (z_open_) inflateEnd(&zs_);
    if (f_) fclose(f_);
  }
  bool failed() const { return failed_; }
  /* next line without its terminator; false at the end of the file or on error (check failed()) */
  bool next(const char*& line, size_t& len) {
    for (;;) {
      const char* base = text_.data() + text_at_;
      const char* nl = (const char*)std::memchr(base, '\n', text_.size() - text_at_);
      if (nl) {
        line = base;
        len = (size_t)(nl - base);
        text_at_ += len + 1;
        if (len && line[len - 1] == '\r') len--;
        return true;
      }
      if (eof_) {
        if (text_at_ < text_.size()) {   // last line without a newline
          line = base;
          len = text_.size() - text_at_;
          text_at_ = text_.size();
          if (len && line[len - 1] == '\r') len--;
          return true;
        }
        return false;
      }
      text_.erase(0, text_at_);   // keep the partial line, fetch more
      text_at_ = 0;
      if (!fill()) return false;
    }
  }

 private:
  bool fill() {
    if (!gz_) {
      if (in_len_ == 0) { eof_ = true; return true; }
      text_.append(in_.data(), in_len_);
      in_len_ = fread(in_.data(), 1, in_.size(), f_);
      if (in_len_ == 0) eof_ = true;
      return true;
    }
    // BGZF = concatenated gzip members
    for (;;) {
      if (zs_.avail_in == 0) {
        const size_t n = fread(in_.data(), 1, in_.size(), f_);
        if (n == 0) {
          if (mid_member_) { err_ = "truncated BGZF stream in " + path_; failed_ = true; return false; }
          eof_ = true;
          return true;
        }
        zs_.next_in = (Bytef*)in_.data();
        zs_.avail_in = (uInt)n;
      }
      zs_.next_out = (Bytef*)out_.data();
      zs_.avail_out = (uInt)out_.size();
      const int rc = inflate(&zs_, Z_NO_FLUSH);
      const size_t got = out_.size() - zs_.avail_out;
      if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) { err_ = "corrupt BGZF stream in " + path_; failed_ = true; return false; }
      mid_member_ = rc != Z_STREAM_END;
      if (rc == Z_STREAM_END && inflateReset(&zs_) != Z_OK) { err_ = "corrupt BGZF stream in " + path_; failed_ = true; return false; }
      if (got) { text_.append(out_.data(), got); return true; }
      if (rc == Z_BUF_ERROR && zs_.avail_in != 0) { err_ = "corrupt BGZF stream in " + path_; failed_ = true; return false; }
    }
  }
  std::string& err_;
  std::string path_, text_;
  size_t text_at_ = 0, in_len_ = 0;
  std::vector<char> in_, out_;
  FILE* f_ = nullptr;
  z_stream zs_;
  bool gz_ = false, z_open_ = false, eof_ = false, failed_ = false, mid_member_ = false;
};

}  // namespace

struct hipstr_snp_vcf {
  std::vector<std::string> samples;
  std::map<std::string, Chrom> chroms;
  std::string sample_text;
  // result of the last region query
  std::vector<int32_t> set_off;
  std::vector<uint32_t> pos;
  std::string base1, base2;
};

namespace {
thread_local std::string g_vcf_error;

bool parse(LineStream& in, hipstr_snp_vcf& v, std::string& err) {
  bool have_header = false;
  const char* line;
  size_t len;
  while (in.next(line, len)) {
    if (len == 0) continue;
    if (line[0] == '#') {
      if (len > 6 && std::memcmp(line, "#CHROM", 6) == 0) {
        std::string header(line, len);
        std::stringstream ss(header);
        std::string item;
        int col = 0;
        while (std::getline(ss, item, '\t'))
          if (col++ >= 9) v.samples.push_back(item);
        have_header = true;
      }
      continue;
    }
    if (!have_header) { err = "VCF record before the #CHROM line"; return false; }
    // split the fixed columns; samples are scanned in place
    const char* end = line + len;
    const char* col[10];
    int n_col = 0;
    const char* p = line;
    col[n_col++] = p;
    while (p < end && n_col < 10) {
      if (*p == '\t') col[n_col++] = p + 1;
      p++;
    }
    if (n_col < 8) { err = "Failed to parse VCF record"; return false; }
    auto width = [&](int c) { return (size_t)((c + 1 < n_col ? col[c + 1] - 1 : end) - col[c]); };
    // biallelic SNP (n_allele == 2 && bcf_is_snp): REF and the single ALT are one character each; like htslib, mpileup's
    // symbolic <X> / <*> alternates count as SNPs too (their "base" is then the '<' the reference takes from allele[0])
    const bool symbolic = width(4) == 3 && col[4][0] == '<' && (col[4][1] == 'X' || col[4][1] == '*') && col[4][2] == '>';
    if (width(3) != 1 || !(symbolic || (width(4) == 1 && col[4][0] != '.'))) continue;
    if (n_col < 10 || v.samples.empty()) continue;
    // index of GT in FORMAT
    int gt_index = -1, k = 0;
    for (const char* q = col[8]; q < col[9] - 1;) {
      const char* stop = (const char*)std::memchr(q, ':', (size_t)(col[9] - 1 - q));
      if (!stop) stop = col[9] - 1;
      if (stop - q == 2 && q[0] == 'G' && q[1] == 'T') { gt_index = k; break; }
      k++;
      q = stop + 1;
    }
    if (gt_index < 0) { err = "Failed to extract the genotypes from the VCF record"; return false; }
    Chrom& c = v.chroms[std::string(col[0], width(0))];
    Site site;
    site.pos = (int32_t)std::strtol(col[1], nullptr, 10);
    site.ref = col[3][0];
    site.alt = col[4][0];
    site.calls = c.codes.size();
    c.codes.resize(c.codes.size() + v.samples.size(), 0);
    const char* q = col[9];
    for (size_t s = 0; s < v.samples.size() && q <= end; s++) {
      const char* stop = (const char*)std::memchr(q, '\t', (size_t)(end - q));
      if (!stop) stop = end;
      // the gt_index-th ':' separated sub-field
      const char* g = q;
      for (int skip = 0; skip < gt_index && g < stop; skip++) {
        const char* colon = (const char*)std::memchr(g, ':', (size_t)(stop - g));
        g = colon ? colon + 1 : stop;
      }
      const char* g_end = (const char*)std::memchr(g, ':', (size_t)(stop - g));
      if (!g_end) g_end = stop;
      // phased diploid call "a|b" with two different single-digit alleles of a biallelic site
      if (g_end - g == 3 && g[1] == '|' && (g[0] == '0' || g[0] == '1') && (g[2] == '0' || g[2] == '1') && g[0] != g[2])
        c.codes[site.calls + s] = g[0] == '0' ? 1 : 2;
      q = stop + 1;
    }
    c.sites.push_back(site);
    c.sorted = false;
  }
  if (in.failed()) return false;
  if (!have_header) { err = "Provided VCF file is improperly formatted"; return false; }
  return true;
}
}  // namespace

extern "C" {

const char* hipstr_snp_vcf_last_error(void) { return g_vcf_error.c_str(); }

hipstr_status_t hipstr_snp_vcf_open(const char* path, hipstr_snp_vcf_t** out) {
  if (!path || !out) return HIPSTR_ERR_BAD_ARG;
  LineStream in(path, g_vcf_error);
  if (in.failed()) return HIPSTR_ERR_BAD_ARG;
  std::unique_ptr<hipstr_snp_vcf> v(new hipstr_snp_vcf());
  if (!parse(in, *v, g_vcf_error)) return HIPSTR_ERR_BAD_ARG;
  for (const std::string& s : v->samples) { v->sample_text += s; v->sample_text += '\n'; }
  *out = v.release();
  return HIPSTR_OK;
}

void hipstr_snp_vcf_close(hipstr_snp_vcf_t* v) { delete v; }

int32_t hipstr_snp_vcf_num_samples(const hipstr_snp_vcf_t* v) { return v ? (int32_t)v->samples.size() : -1; }
const char* hipstr_snp_vcf_samples(const hipstr_snp_vcf_t* v) { return v ? v->sample_text.c_str() : nullptr; }
int32_t hipstr_snp_vcf_has_chromosome(const hipstr_snp_vcf_t* v, const char* chrom) {
  return v && chrom && v->chroms.count(chrom) ? 1 : 0;
}

hipstr_status_t hipstr_snp_vcf_region_sets(hipstr_snp_vcf_t* v, const char* chrom, int32_t start, int32_t end, int32_t n_skip,
                                           const int32_t* skip_start, const int32_t* skip_stop, int32_t skip_padding, int32_t* found,
                                           const int32_t** set_off, const uint32_t** snp_pos, const char** snp_base1,
                                           const char** snp_base2) {
  if (!v || !chrom || !found || !set_off || !snp_pos || !snp_base1 || !snp_base2 || n_skip < 0 || (n_skip > 0 && (!skip_start || !skip_stop)))
    return HIPSTR_ERR_BAD_ARG;
  const size_t S = v->samples.size();
  v->set_off.assign(S + 1, 0);
  v->pos.clear(); v->base1.clear(); v->base2.clear();
  *set_off = v->set_off.data();
  auto it = v->chroms.find(chrom);
  *found = it != v->chroms.end();      // VCFReader::set_region fails for a chromosome the index does not know
  if (it != v->chroms.end()) {
    Chrom& c = it->second;
    if (!c.sorted) {
      std::stable_sort(c.sites.begin(), c.sites.end(), [](const Site& a, const Site& b) { return a.pos < b.pos; });
      c.sorted = true;
    }
    auto lo = std::lower_bound(c.sites.begin(), c.sites.end(), start, [](const Site& s, int32_t p) { return s.pos < p; });
    auto hi = std::upper_bound(c.sites.begin(), c.sites.end(), end, [](int32_t p, const Site& s) { return p < s.pos; });
    std::vector<const Site*> kept;
    for (auto s = lo; s != hi; ++s) {
      bool skip = false;
      for (int r = 0; r < n_skip && !skip; r++) skip = s->pos >= skip_start[r] - skip_padding && s->pos <= skip_stop[r] + skip_padding;
      if (!skip) kept.push_back(&*s);
    }
    for (size_t smp = 0; smp < S; smp++) {
      for (const Site* s : kept) {
        const uint8_t code = c.codes[s->calls + smp];
        if (!code) continue;
        v->pos.push_back((uint32_t)(s->pos - 1));       // VCF is 1-based, the alignments 0-based
        v->base1 += code == 1 ? s->ref : s->alt;
        v->base2 += code == 1 ? s->alt : s->ref;
      }
      v->set_off[smp + 1] = (int32_t)v->pos.size();
    }
  }
  if (v->pos.empty()) v->pos.push_back(0);
  *snp_pos = v->pos.data();
  *snp_base1 = v->base1.c_str();
  *snp_base2 = v->base2.c_str();
  return HIPSTR_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// The reference panel (--ref-vcf): the STR record of a region, read_vcf_alleles (src/vcf_input.cpp:21-50).
// ---------------------------------------------------------------------------------------------
struct hipstr_str_vcf {
  struct Record { int32_t pos; int32_t ref_len; bool has_span; int32_t start, end; std::vector<std::string> alleles; };
  std::map<std::string, std::vector<Record> > chroms;   // file order
  std::string result;
};

namespace {
bool info_int(const std::string& info, const char* key, int32_t& value) {   // KEY=<int> among the ';' separated INFO entries
  const size_t klen = std::strlen(key);
  size_t at = 0;
  while (at <= info.size()) {
    size_t stop = info.find(';', at);
    if (stop == std::string::npos) stop = info.size();
    if (stop - at > klen && info.compare(at, klen, key) == 0 && info[at + klen] == '=') {
      value = (int32_t)std::strtol(info.c_str() + at + klen + 1, nullptr, 10);
      return true;
    }
    at = stop + 1;
  }
  return false;
}
}  // namespace

extern "C" {

hipstr_status_t hipstr_str_vcf_open(const char* path, hipstr_str_vcf_t** out) {
  if (!path || !out) return HIPSTR_ERR_BAD_ARG;
  LineStream in(path, g_vcf_error);
  if (in.failed()) return HIPSTR_ERR_BAD_ARG;
  std::unique_ptr<hipstr_str_vcf> v(new hipstr_str_vcf());
  const char* raw;
  size_t raw_len;
  while (in.next(raw, raw_len)) {
    const std::string line(raw, raw_len);
    if (line.empty() || line[0] == '#') continue;
    std::vector<std::string> f;
    size_t at = 0;
    for (int c = 0; c < 8; c++) {
      const size_t tab = line.find('\t', at);
      f.push_back(line.substr(at, tab == std::string::npos ? std::string::npos : tab - at));
      if (tab == std::string::npos) break;
      at = tab + 1;
    }
    if (f.size() < 8) { g_vcf_error = "Failed to parse VCF record"; return HIPSTR_ERR_BAD_ARG; }
    hipstr_str_vcf::Record r;
    r.pos = (int32_t)std::strtol(f[1].c_str(), nullptr, 10);
    r.ref_len = (int32_t)f[3].size();
    r.alleles.push_back(f[3]);
    if (f[4] != ".") {
      std::stringstream alts(f[4]);
      std::string a;
      while (std::getline(alts, a, ',')) r.alleles.push_back(a);
    }
    const bool has_start = info_int(f[7], "START", r.start), has_end = info_int(f[7], "END", r.end);
    r.has_span = has_start && has_end;
    v->chroms[f[0]].push_back(r);
  }
  if (in.failed()) return HIPSTR_ERR_BAD_ARG;
  *out = v.release();
  return HIPSTR_OK;
}

void hipstr_str_vcf_close(hipstr_str_vcf_t* v) { delete v; }

int32_t hipstr_str_vcf_alleles(hipstr_str_vcf_t* v, const char* chrom, int32_t region_start, int32_t region_stop, int32_t* pos,
                               int32_t* n_alleles, const char** alleles_text) {
  if (!v || !chrom || !pos || !n_alleles || !alleles_text) return -1;
  *pos = -1;
  *n_alleles = 0;
  v->result.clear();
  *alleles_text = v->result.c_str();
  auto it = v->chroms.find(chrom);
  if (it == v->chroms.end()) return 0;
  const int32_t pad = 50;   // vcf_input.cpp:19
  const int32_t pad_start = region_start < pad ? 0 : region_start - pad;
  const int32_t beg = pad_start > 0 ? pad_start - 1 : 0, end = region_stop + pad;   // the tabix region "chrom:pad_start-end"
  for (const hipstr_str_vcf::Record& r : it->second) {
    if (!(r.pos - 1 < end && r.pos - 1 + r.ref_len > beg)) continue;
    if (!r.has_span) continue;          // not an STR record
    if (r.start == region_start + 1 && r.end == region_stop) {
      *pos = r.pos - 1;
      *n_alleles = (int32_t)r.alleles.size();
      for (const std::string& a : r.alleles) { v->result += a; v->result += '\n'; }
      *alleles_text = v->result.c_str();
      return 1;
    }
    if (r.pos > region_start + pad) break;
  }
  return 0;
}

}  // extern "C"
