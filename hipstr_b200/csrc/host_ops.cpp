/*
 * host_ops.cpp -- the integer / byte host logic of the hot path that stays on
 * the CPU: seed selection (a5) and read pooling (a2).  Pure C++, no CUDA.
 * Reference behaviour: SeqAlignment/HapAligner.cpp:238-318, read_pooler.cpp:3-20,
 * base_quality.cpp:11-28.
 */
#include <algorithm>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/hipstr_b200.h"

namespace {

const int kMinSeedDist = 5;  // HapAligner.cpp:17

// Best seed inside [lo, hi] (genomic, inclusive): the centre of the widest
// stretch not covered by a repeat block; the later stretch wins ties and
// within a stretch of even length the left of the two central bases is used
// (calc_best_seed_position, HapAligner.cpp:238-264).
inline void widest_free_stretch(int32_t lo, int32_t hi, int32_t n_rep, const int32_t* rs, const int32_t* re,
                                int32_t& dist, int32_t& pos) {
  dist = pos = -1;
  int32_t cursor = lo;
  auto offer = [&](int32_t a, int32_t b) {  // free stretch [a, b]
    int32_t d = 1 + (b - a) / 2;
    if (d >= dist) { dist = d; pos = a + d - 1; }
  };
  for (int32_t i = 0; i < n_rep && cursor <= hi; i++) {
    if (cursor < rs[i]) {
      offer(cursor, std::min(hi, rs[i] - 1));
      cursor = re[i];
    } else if (cursor < re[i])
      cursor = re[i];
  }
  if (cursor <= hi) offer(cursor, hi);
}

}  // namespace

extern "C" hipstr_status_t hipstr_calc_seeds(int32_t n_reads, const int32_t* read_start, const int32_t* read_len,
                                             const int32_t* cigar_off, const char* cigar_type,
                                             const int32_t* cigar_len, int32_t first_block_start,
                                             int32_t last_block_end, int32_t n_repeats,
                                             const int32_t* repeat_start, const int32_t* repeat_end,
                                             int32_t* out_seed) {
  if (n_reads < 0 || (n_reads > 0 && (!read_start || !read_len || !cigar_off || !out_seed))) return HIPSTR_ERR_BAD_ARG;
  for (int32_t r = 0; r < n_reads; r++) {
    int32_t gpos = read_start[r];  // genomic coordinate of the next reference-consuming base
    int32_t rpos = 0;              // index of the next read base
    int32_t seed = -1, bar = kMinSeedDist;
    for (int32_t c = cigar_off[r]; c < cigar_off[r + 1]; c++) {
      const int32_t n = cigar_len[c];
      const char t = cigar_type[c];
      if (t == '=') {
        int32_t lo = std::max(gpos, first_block_start);
        int32_t hi = std::min(gpos + n - 1, last_block_end - 1);
        if (lo <= hi) {
          int32_t d, p;
          widest_free_stretch(lo, hi, n_repeats, repeat_start, repeat_end, d, p);
          if (d >= bar) { bar = d; seed = rpos + (p - gpos); }
        }
        gpos += n; rpos += n;
      } else if (t == 'X') {
        gpos += n; rpos += n;
      } else if (t == 'I') {
        rpos += n;
      } else if (t == 'D') {
        gpos += n;
      } else
        return HIPSTR_ERR_BAD_CIGAR;
    }
    if (seed < -1 || seed == 0 || seed >= read_len[r] - 1) return HIPSTR_ERR_INVALID_SEED;
    out_seed[r] = seed;
  }
  return HIPSTR_OK;
}

extern "C" hipstr_status_t hipstr_pool_reads(int32_t n_reads, const int32_t* seq_off, const char* bases,
                                             const char* quals, int32_t* pool_index, int32_t* n_pools,
                                             int32_t* pool_first_read, int32_t* pool_seq_off, char* pool_bases,
                                             char* pool_quals) {
  if (n_reads < 0 || !n_pools) return HIPSTR_ERR_BAD_ARG;
  if (n_reads > 0 && (!seq_off || !bases || !quals || !pool_index || !pool_first_read || !pool_seq_off ||
                      !pool_bases || !pool_quals))
    return HIPSTR_ERR_BAD_ARG;
  // keys are views into the caller's bases: no per-read string allocations
  std::unordered_map<std::string_view, int32_t> by_seq;
  by_seq.reserve((size_t)n_reads * 2);
  std::vector<int32_t> next_member((size_t)n_reads, -1), last_member, count;   // members of a pool as a linked list in read order
  for (int32_t r = 0; r < n_reads; r++) {
    const std::string_view key(bases + seq_off[r], (size_t)(seq_off[r + 1] - seq_off[r]));
    auto it = by_seq.find(key);
    if (it == by_seq.end()) {
      it = by_seq.emplace(key, (int32_t)last_member.size()).first;
      pool_first_read[last_member.size()] = r;
      last_member.push_back(r);
      count.push_back(1);
    } else {
      next_member[last_member[it->second]] = r;
      last_member[it->second] = r;
      count[it->second]++;
    }
    pool_index[r] = it->second;
  }
  int32_t at = 0;
  std::vector<const unsigned char*> rows;
  std::vector<unsigned char> work;
  uint32_t hist[256];   // a pool can hold more than 65 535 reads (amplicon data)
  std::memset(hist, 0, sizeof(hist));
  for (size_t p = 0; p < last_member.size(); p++) {
    const int32_t first = pool_first_read[p];
    const int32_t len = seq_off[first + 1] - seq_off[first];
    pool_seq_off[p] = at;
    std::memcpy(pool_bases + at, bases + seq_off[first], len);
    const size_t m = (size_t)count[p];
    if (m == 1)
      std::memcpy(pool_quals + at, quals + seq_off[first], len);
    else if (m == 2) {
      // upper median of two = the larger byte in signed char order (std::sort on chars, base_quality.cpp:11-28)
      const signed char* a = reinterpret_cast<const signed char*>(quals + seq_off[first]);
      const signed char* b = reinterpret_cast<const signed char*>(quals + seq_off[next_member[first]]);
      for (int32_t i = 0; i < len; i++) pool_quals[at + i] = (char)(a[i] > b[i] ? a[i] : b[i]);
    } else if (m <= 24) {
      // Upper median per position (sorted[m / 2] in signed char order) of a few rows: an odd-even transposition network
      // over whole rows -- every compare-exchange is an elementwise min / max of two byte rows, which the compiler turns
      // into vector instructions.  (byte ^ 0x80) turns signed order into unsigned order.
      const size_t pitch = ((size_t)len + 63) & ~(size_t)63;
      work.resize(m * pitch);
      size_t k = 0;
      for (int32_t r = first; r >= 0; r = next_member[r], k++) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(quals + seq_off[r]);
        unsigned char* dst = work.data() + k * pitch;
        for (int32_t i = 0; i < len; i++) dst[i] = src[i] ^ 0x80u;
      }
      for (size_t pass = 0; pass < m; pass++)
        for (size_t a = pass & 1; a + 1 < m; a += 2) {
          unsigned char* __restrict lo = work.data() + a * pitch;
          unsigned char* __restrict hi = lo + pitch;
          for (int32_t i = 0; i < len; i++) {
            const unsigned char x = lo[i], y = hi[i];
            lo[i] = x < y ? x : y;
            hi[i] = x < y ? y : x;
          }
        }
      const unsigned char* med = work.data() + (m / 2) * pitch;
      for (int32_t i = 0; i < len; i++) pool_quals[at + i] = (char)(med[i] ^ 0x80u);
    } else {
      // Upper median per position (sorted[m / 2] in signed char order) by counting: m increments of a 256-bin histogram
      // indexed by (byte ^ 0x80) -- which turns signed order into unsigned order -- then a scan of the occupied range.
      rows.clear();
      for (int32_t r = first; r >= 0; r = next_member[r]) rows.push_back(reinterpret_cast<const unsigned char*>(quals + seq_off[r]));
      const uint32_t want = (uint32_t)(m / 2);
      for (int32_t i = 0; i < len; i++) {
        unsigned lo = 255, hi = 0;
        for (size_t k = 0; k < m; k++) {
          const unsigned v = rows[k][i] ^ 0x80u;
          hist[v]++;
          lo = v < lo ? v : lo;
          hi = v > hi ? v : hi;
        }
        uint32_t seen = 0;
        unsigned med = hi;
        bool found = false;
        for (unsigned v = lo; v <= hi; v++) {
          seen += hist[v];
          if (!found && seen > want) { med = v; found = true; }
          hist[v] = 0;
        }
        pool_quals[at + i] = (char)(med ^ 0x80u);
      }
    }
    at += len;
  }
  const std::vector<int32_t>& members = last_member;
  pool_seq_off[members.size()] = at;
  *n_pools = (int32_t)members.size();
  return HIPSTR_OK;
}

// ---------------------------------------------------------------------------------------------
// Trace -> alignment against the reference (AlignmentTraceback.cpp:5-144).
// ---------------------------------------------------------------------------------------------
namespace {

// Composes the two op strings walking away from the seed column in direction `step`
// (stitch, AlignmentTraceback.cpp:5-53).  Returns false on an inconsistent pair.
bool compose(const std::string& hap, const std::string& read, long h, long r, int step, std::string& out) {
  out.clear();
  while (r >= 0 && r < (long)read.size()) {
    const char rc = read[r];
    if (rc == 'S') { out += 'S'; r += step; continue; }
    if (h < 0 || h >= (long)hap.size()) return false;
    const char hc = hap[h];
    if (hc == 'D') {                      // reference base absent from the haplotype
      if (rc == 'I') { out += 'M'; r += step; h += step; }   // ... but the read re-inserts a base there
      else { out += 'D'; h += step; }
    } else if (rc == 'I') { out += 'I'; r += step; }
    else if (rc == 'D') {
      if (hc == 'M') out += 'D';
      else if (hc != 'I') return false;   // haplotype insertion that the read deletes: nothing to emit
      r += step; h += step;
    } else if (rc == 'M') {
      if (hc != 'M' && hc != 'I') return false;
      out += hc; r += step; h += step;
    } else
      return false;
  }
  return true;
}

// The same walk when only the reference span is wanted: the number of reference-consuming operations ('M' / 'D') it
// would emit, no string built.  Returns -1 on an inconsistent pair.
long compose_span(const char* hap, long hap_n, const char* read, long read_n, long h, long r, int step) {
  long span = 0;
  while (r >= 0 && r < read_n) {
    const char rc = read[r];
    if (rc == 'S') { r += step; continue; }
    if (h < 0 || h >= hap_n) return -1;
    const char hc = hap[h];
    if (hc == 'D') {
      span++;
      if (rc == 'I') r += step;
      h += step;
    } else if (rc == 'I') r += step;
    else if (rc == 'D') {
      if (hc == 'M') span++;
      else if (hc != 'I') return -1;
      r += step; h += step;
    } else if (rc == 'M') {
      if (hc == 'M') span++;
      else if (hc != 'I') return -1;
      r += step; h += step;
    } else
      return -1;
  }
  return span;
}

}  // namespace

extern "C" hipstr_status_t hipstr_hap_aln_index(const char* hap, int32_t n, int32_t* index) {
  if (!hap || n < 0 || !index) return HIPSTR_ERR_BAD_ARG;
  int32_t *non_d = index, *non_i = index + (n + 1), *col_of = index + 2 * (n + 1);
  non_d[0] = non_i[0] = 0;
  for (int32_t c = 0; c < n; c++) {
    if (hap[c] != 'M' && hap[c] != 'I' && hap[c] != 'D') return HIPSTR_ERR_BAD_ARG;
    if (hap[c] != 'D') col_of[non_d[c]] = c;
    non_d[c + 1] = non_d[c] + (hap[c] != 'D');
    non_i[c + 1] = non_i[c] + (hap[c] != 'I');
  }
  for (int32_t k = non_d[n]; k <= n; k++) col_of[k] = n;   // "no such operation"
  return HIPSTR_OK;
}

extern "C" hipstr_status_t hipstr_trace_span(int32_t hap_start, const char* hap, int32_t hap_n, const int32_t* index,
                                             const char* read, int32_t seed_hap_pos, int32_t seed_base, int32_t* start,
                                             int32_t* stop) {
  if (!hap || !index || !read || !start || !stop || hap_n < 0) return HIPSTR_ERR_BAD_ARG;
  const int32_t *non_d = index, *non_i = index + (hap_n + 1), *col_of = index + 2 * (hap_n + 1);
  // the seed's column in the haplotype string: after seed_hap_pos non-'D' operations, then over any 'D's
  if (seed_hap_pos < 0 || seed_hap_pos >= non_d[hap_n]) {
    // fewer haplotype bases than the seed position (or a negative one): let the stepping form decide
    int32_t n_cigar = 0;
    return hipstr_stitch_trace(hap_start, hap, read, seed_hap_pos, seed_base, "", start, stop, 0, nullptr, nullptr, &n_cigar, 0, nullptr);
  }
  const long first = seed_hap_pos == 0 ? 0 : col_of[seed_hap_pos - 1] + 1;   // where the counting loop of the stepping form stops
  const int32_t seed_pos = hap_start + non_i[first];
  const long hcol = col_of[seed_hap_pos];
  // one pass over the read's operations: the seed's column, the outermost aligned ('M' / 'D') operation on either side and
  // how many there are; anything but M, I, D, S goes to the stepping form
  long read_n = 0, rcol = -1, remaining = seed_base, r_lo = -1, r_hi = -1, n_left = 0, n_right = 0;
  bool odd = false;
  for (; read[read_n]; read_n++) {
    const char c = read[read_n];
    odd |= !(c == 'M' || c == 'I' || c == 'D' || c == 'S');
    if (rcol < 0) {
      if (remaining > 0) remaining -= c != 'D';
      else if (c != 'D') rcol = read_n;
    }
  }
  if (rcol < 0 || odd || seed_base < 0) {
    int32_t n_cigar = 0;
    return hipstr_stitch_trace(hap_start, hap, read, seed_hap_pos, seed_base, "", start, stop, 0, nullptr, nullptr, &n_cigar, 0, nullptr);
  }
  for (long r = 0; r < rcol; r++)
    if (read[r] == 'M' || read[r] == 'D') { if (r_lo < 0) r_lo = r; n_left++; }
  for (long r = rcol + 1; r < read_n; r++)
    if (read[r] == 'M' || read[r] == 'D') { r_hi = r; n_right++; }
  // left of the seed: the n_left aligned operations consume the n_left non-'D' columns before the seed's and every 'D' between
  long span_left = 0, h = hcol - 1, r = rcol - 1;
  if (n_left > 0) {
    if (n_left > non_d[hcol]) return HIPSTR_ERR_BAD_ARG;
    const long p = col_of[non_d[hcol] - n_left];
    span_left = non_i[hcol] - non_i[p];
    h = p - 1;
    r = r_lo - 1;
  }
  const long tail_left = compose_span(hap, hap_n, read, read_n, h, r, -1);
  long span_right = 0;
  h = hcol + 1; r = rcol + 1;
  if (n_right > 0) {
    if (n_right > non_d[hap_n] - non_d[hcol + 1]) return HIPSTR_ERR_BAD_ARG;
    const long p = col_of[non_d[hcol + 1] + n_right - 1];
    span_right = non_i[p + 1] - non_i[hcol + 1];
    h = p + 1;
    r = r_hi + 1;
  }
  const long tail_right = compose_span(hap, hap_n, read, read_n, h, r, 1);
  if (tail_left < 0 || tail_right < 0) return HIPSTR_ERR_BAD_ARG;
  *start = seed_pos - (int32_t)(span_left + tail_left);
  *stop = seed_pos + (int32_t)(span_right + tail_right);
  return HIPSTR_OK;
}

extern "C" hipstr_status_t hipstr_stitch_trace(int32_t hap_start, const char* hap_aln_to_ref, const char* read_aln_to_hap,
                                               int32_t seed_hap_pos, int32_t seed_base, const char* read_bases,
                                               int32_t* start, int32_t* stop, int32_t cigar_cap, char* cigar_type,
                                               int32_t* cigar_len, int32_t* n_cigar, int32_t aln_cap, char* alignment) {
  if (!hap_aln_to_ref || !read_aln_to_hap || !read_bases || !start || !stop || !n_cigar) return HIPSTR_ERR_BAD_ARG;
  const bool want_strings = cigar_type && cigar_len && alignment;   // NULL = only the span is wanted
  if (!want_strings) {
    const char *hap = hap_aln_to_ref, *read = read_aln_to_hap;
    const long hap_n = (long)std::strlen(hap), read_n = (long)std::strlen(read);
    long hcol = 0, remaining = seed_hap_pos;
    int32_t seed_pos = hap_start;
    for (; remaining > 0 && hcol < hap_n; hcol++) {
      remaining -= hap[hcol] != 'D';
      seed_pos += hap[hcol] != 'I';
    }
    while (hcol < hap_n && hap[hcol] == 'D') hcol++;
    if (hcol == hap_n) return HIPSTR_ERR_BAD_ARG;
    long rcol = 0;
    for (remaining = seed_base; remaining > 0 && rcol < read_n; rcol++) remaining -= read[rcol] != 'D';
    while (rcol < read_n && read[rcol] == 'D') rcol++;
    if (rcol == read_n) return HIPSTR_ERR_BAD_ARG;
    const long left = compose_span(hap, hap_n, read, read_n, hcol - 1, rcol - 1, -1);
    const long right = compose_span(hap, hap_n, read, read_n, hcol + 1, rcol + 1, 1);
    if (left < 0 || right < 0) return HIPSTR_ERR_BAD_ARG;
    *start = seed_pos - (int32_t)left;
    *stop = seed_pos + (int32_t)right;
    *n_cigar = 0;
    return HIPSTR_OK;
  }
  const std::string hap(hap_aln_to_ref), read(read_aln_to_hap);
  // column of the haplotype-vs-reference alignment that holds haplotype base seed_hap_pos, and its coordinate
  long hcol = 0, remaining = seed_hap_pos;
  int32_t seed_pos = hap_start;
  while (remaining > 0 && hcol < (long)hap.size()) {
    if (hap[hcol] == 'M' || hap[hcol] == 'I') remaining--;
    if (hap[hcol] == 'M' || hap[hcol] == 'D') seed_pos++;
    hcol++;
  }
  while (hcol < (long)hap.size() && hap[hcol] == 'D') hcol++;
  if (hcol == (long)hap.size()) return HIPSTR_ERR_BAD_ARG;
  // column of the read-vs-haplotype string that holds the seed base
  long rcol = 0;
  remaining = seed_base;
  while (remaining > 0 && rcol < (long)read.size()) {
    if (read[rcol] == 'M' || read[rcol] == 'I' || read[rcol] == 'S') remaining--;
    rcol++;
  }
  while (rcol < (long)read.size() && read[rcol] == 'D') rcol++;
  if (rcol == (long)read.size()) return HIPSTR_ERR_BAD_ARG;
  std::string left, right;
  if (!compose(hap, read, hcol - 1, rcol - 1, -1, left) || !compose(hap, read, hcol + 1, rcol + 1, 1, right)) return HIPSTR_ERR_BAD_ARG;
  std::reverse(left.begin(), left.end());
  std::string full = left + "M" + right;
  for (size_t i = 0; i < full.size() && full[i] == 'I'; i++) full[i] = 'S';   // leading insertion = soft clip
  int32_t a = seed_pos, b = seed_pos;
  for (char c : left) if (c == 'D' || c == 'M') a--;
  for (char c : right) if (c == 'D' || c == 'M') b++;
  *start = a;
  *stop = b;
  *n_cigar = 0;
  if (!want_strings) return HIPSTR_OK;
  int32_t runs = 0;
  for (size_t i = 0; i < full.size();) {
    size_t k = i;
    while (k < full.size() && full[k] == full[i]) k++;
    if (runs >= cigar_cap) return HIPSTR_ERR_BAD_ARG;
    cigar_type[runs] = full[i];
    cigar_len[runs] = (int32_t)(k - i);
    runs++;
    i = k;
  }
  *n_cigar = runs;
  const size_t n_bases = std::strlen(read_bases);
  size_t ri = 0, out = 0;
  for (char c : full) {
    if (c == 'S') { ri++; continue; }
    if (out + 1 >= (size_t)aln_cap) return HIPSTR_ERR_BAD_ARG;
    if (c == 'D') alignment[out++] = '-';
    else {
      if (ri >= n_bases) return HIPSTR_ERR_BAD_ARG;
      alignment[out++] = read_bases[ri++];
    }
  }
  alignment[out] = 0;
  return HIPSTR_OK;
}
