/*
 * left_align.cpp -- SURVEY.md 8(f) row 3: GenotyperBamProcessor::left_align_reads (src/genotyper_bam_processor.cpp:38-102)
 * for a batch of loci.  Per read the reference trims the BAM alignment to the region +- 40 bp (BamAlignment::TrimAlignment,
 * src/bam_io.cpp:384-477), then either re-expresses an indel-free alignment with =/X operations (convertAlignment,
 * src/SeqAlignment/AlignmentOps.cpp:102-167) or re-aligns the read against a window of the chromosome with
 * Needleman-Wunsch (realign, :14-100), once per distinct sequence.  Here the Needleman-Wunsch alignments of ALL loci go
 * to the GPU in one call (K6, hipstr_nw_align_batch_host); the bookkeeping around them is the reference's, on the host.
 * The result is a hipstr_locus_reads_t -- exactly what hipstr_genotyper_create_from_reads consumes.
 */
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "seq_stutter_genotyper.h"

namespace hipstr {

namespace {

const int kAlignWindowWidth = 75;   // ALIGN_WINDOW_WIDTH, AlignmentOps.cpp:8

typedef std::vector<std::pair<char, int32_t> > Cigar;

struct Trimmed {          // a BamAlignment after TrimAlignment
  int32_t pos, end_pos;   // Position(), GetEndPosition() (exclusive)
  std::string bases, quals;
  Cigar cigar;
};

struct Aligned {          // the fields of the reference's Alignment that travel on
  int32_t start = 0, stop = 0;
  std::string bases, quals;
  Cigar cigar;
  bool ok = true;
};

inline char up(char c) { return (char)std::toupper((unsigned char)c); }

/* BamAlignment::TrimAlignment with the default quality bound '~': bases outside [min_read_start, max_read_stop] are cut
 * off operation by operation from both ends. */
void trim_alignment(Trimmed& a, int32_t min_read_start, int32_t max_read_stop, char min_base_qual = '~') {
  int ltrim = 0;
  int32_t start_pos = a.pos;
  size_t front = 0;
  while (start_pos < min_read_start && front < a.cigar.size()) {
    const char t = a.cigar[front].first;
    if ((t == 'M' || t == '=' || t == 'X' || t == 'I' || t == 'S') && a.quals[ltrim] > min_base_qual) break;
    if (t == 'M' || t == '=' || t == 'X') { ltrim++; start_pos++; }
    else if (t == 'D') start_pos++;
    else if (t == 'I' || t == 'S') ltrim++;
    if (a.cigar[front].second == 1) front++;
    else a.cigar[front].second--;
  }
  a.cigar.erase(a.cigar.begin(), a.cigar.begin() + front);
  int rtrim = 0;
  const int last = (int)a.quals.size() - 1;
  int32_t end_pos = a.end_pos;
  while (end_pos > max_read_stop && !a.cigar.empty()) {
    const char t = a.cigar.back().first;
    if ((t == 'M' || t == '=' || t == 'X' || t == 'I' || t == 'S') && a.quals[last - rtrim] > min_base_qual) break;
    if (t == 'M' || t == '=' || t == 'X') { rtrim++; end_pos--; }
    else if (t == 'D') end_pos--;
    else if (t == 'I' || t == 'S') rtrim++;
    if (a.cigar.back().second == 1) a.cigar.pop_back();
    else a.cigar.back().second--;
  }
  a.bases = a.bases.substr(ltrim, a.bases.size() - ltrim - rtrim);
  a.quals = a.quals.substr(ltrim, a.quals.size() - ltrim - rtrim);
  a.pos = start_pos;
  a.end_pos = end_pos;
}

bool matches_reference(const Trimmed& a) {   // BamAlignment::MatchesReference, bam_io.h:244-250
  for (const auto& op : a.cigar)
    if (op.first != 'M' && op.first != '=') return false;
  return true;
}

/* convertAlignment: M runs are split into = / X by comparing with the chromosome; runs do not merge across operations */
Aligned convert_alignment(const Trimmed& a, const char* chrom) {
  Aligned out;
  out.start = a.pos;
  out.stop = a.end_pos - 1;
  out.quals = a.quals;
  out.bases = a.bases;
  for (char& c : out.bases) c = up(c);
  int32_t seq_index = 0, ref_index = a.pos;
  for (const auto& op : a.cigar) {
    switch (op.first) {
      case 'H': break;
      case 'S': case 'I': out.cigar.push_back(op); seq_index += op.second; break;
      case 'D': out.cigar.push_back(op); ref_index += op.second; break;
      default: {   // M, =, X
        char run_type = '=';
        int32_t run = 0;
        for (int32_t k = 0; k < op.second; k++, ref_index++, seq_index++) {
          const char want = out.bases[seq_index] == up(chrom[ref_index]) ? '=' : 'X';
          if (want == run_type) run++;
          else {
            if (run != 0) out.cigar.emplace_back(run_type, run);
            run_type = want;
            run = 1;
          }
        }
        if (run != 0) out.cigar.emplace_back(run_type, run);
      }
    }
  }
  return out;
}

inline int base_class(char c) {   // NeedlemanWunsch.cpp:105-123
  switch (up(c)) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

/* The tail of realign() (AlignmentOps.cpp:33-99) on the operation string K6 returns for (window, read). */
Aligned finish_realign(const Trimmed& a, int32_t window_start, const char* window, const char* ops, int32_t n_ops) {
  Aligned out;
  if (n_ops < 0) { out.ok = false; return out; }
  // columns before / after the aligned part are reference bases against gaps
  int lead = 0, trail = 0;
  while (lead < n_ops && ops[lead] == 'D') lead++;
  while (trail < n_ops - lead && ops[n_ops - 1 - trail] == 'D') trail++;
  // the reference's CIGAR: = / X by base class, D, I, run-length encoded (traceAlignment :269-337)
  Cigar cigar;
  int32_t ref_at = lead, read_at = 0, ref_span = 0;
  for (int k = lead; k < n_ops - trail; k++) {
    char t;
    if (ops[k] == 'M') { t = base_class(window[ref_at]) == base_class(a.bases[read_at]) ? '=' : 'X'; ref_at++; read_at++; ref_span++; }
    else if (ops[k] == 'D') { t = 'D'; ref_at++; ref_span++; }
    else { t = 'I'; read_at++; }
    if (!cigar.empty() && cigar.back().first == t) cigar.back().second++;
    else cigar.emplace_back(t, 1);
  }
  out.start = window_start + lead;
  out.stop = out.start + ref_span - 1;
  // read bases that hang over the window's ends (insertions at the very ends) are clipped away
  int head = 0, back = 0;
  while (head < n_ops && ops[head] == 'I') head++;
  while (back < n_ops - 1 && ops[n_ops - 1 - back] == 'I') back++;
  const int n = (int)a.bases.size();
  out.quals = a.quals.substr(head, n - head - back);
  out.bases = a.bases.substr(head, n - head - back);
  for (char& c : out.bases) c = up(c);
  if (cigar.empty()) { out.ok = false; return out; }
  int h = head, t = back;
  size_t end = cigar.size() - 1;
  while (t > cigar[end].second && end != 0) { t -= cigar[end].second; end--; }
  for (size_t k = 0; k < end; k++) {
    if (h >= cigar[k].second) h -= cigar[k].second;
    else if (h > 0) { out.cigar.emplace_back(cigar[k].first, cigar[k].second - h); h = 0; }
    else out.cigar.push_back(cigar[k]);
  }
  if (h + t > cigar[end].second) { out.ok = false; return out; }   // the reference dies here
  if (h + t < cigar[end].second) out.cigar.emplace_back(cigar[end].first, cigar[end].second - h - t);
  return out;
}

}  // namespace
}  // namespace hipstr

struct hipstr_left_aligned {
  hipstr_locus_reads_t view;
  std::vector<int32_t> locus_read_off, locus_sample_off, read_seq_off, read_start, read_stop, cigar_off, cigar_len, sample_label, name_id,
      source;
  std::vector<char> bases, quals, cigar_type;
  std::vector<double> log_p1, log_p2;
  std::vector<uint8_t> haploid, rev_strand, use_for_haps;
  int64_t fail_count = 0, nw_alignments = 0;
};

extern "C" {

hipstr_status_t hipstr_left_align_reads_host(hipstr_ctx_t* ctx, int32_t n_loci, const hipstr_locus_reads_t* raw,
                                             const char* const* chrom_seq, const int32_t* trim_start, const int32_t* trim_stop,
                                             hipstr_left_aligned_t** out_handle) {
  using namespace hipstr;
  if (!ctx) return HIPSTR_ERR_NO_DEVICE;   // the alignments run on the GPU; there is no CPU path
  if (!raw || !chrom_seq || !out_handle || n_loci < 0 || !raw->read_stop) return HIPSTR_ERR_BAD_ARG;
  const int R = raw->locus_read_off[n_loci];
  std::vector<Trimmed> reads(R);
  std::vector<int> locus_of(R);
  std::vector<size_t> chrom_len(n_loci);
  {
    // one strlen per distinct chromosome of the window, not one per locus (a chromosome is up to 250 MB)
    std::map<const char*, size_t> length_of;
    for (int l = 0; l < n_loci; l++) {
      if (!chrom_seq[l]) return HIPSTR_ERR_BAD_ARG;
      auto it = length_of.find(chrom_seq[l]);
      if (it == length_of.end()) it = length_of.emplace(chrom_seq[l], std::strlen(chrom_seq[l])).first;
      chrom_len[l] = it->second;
    }
  }
  // every step below is per locus and independent: the loci of the window are spread over the host threads
  parallel_for((size_t)n_loci, [&](size_t li) {
    const int l = (int)li;
    for (int r = raw->locus_read_off[l]; r < raw->locus_read_off[l + 1]; r++) {
      Trimmed& a = reads[r];
      locus_of[r] = l;
      a.pos = raw->read_start[r];
      a.end_pos = raw->read_stop[r];
      a.bases.assign(raw->bases + raw->read_seq_off[r], raw->bases + raw->read_seq_off[r + 1]);
      a.quals.assign(raw->quals + raw->read_seq_off[r], raw->quals + raw->read_seq_off[r + 1]);
      for (int c = raw->cigar_off[r]; c < raw->cigar_off[r + 1]; c++) a.cigar.emplace_back(raw->cigar_type[c], raw->cigar_len[c]);
      if (trim_start && trim_stop) trim_alignment(a, trim_start[l], trim_stop[l]);
    }
  });
  // which reads need a Needleman-Wunsch alignment: the first occurrence of every distinct sequence of a locus unless
  // its CIGAR is indel- and clip-free (round 1); later occurrences only when that first result was clipped (round 2)
  std::vector<std::map<std::string, std::vector<int> > > occurrences(n_loci);
  parallel_for((size_t)n_loci, [&](size_t l) {
    for (int r = raw->locus_read_off[l]; r < raw->locus_read_off[l + 1]; r++)
      if (!reads[r].bases.empty()) occurrences[l][reads[r].bases].push_back(r);
  });
  std::vector<Aligned> realigned(R);          // read index -> result of realign()
  std::vector<uint8_t> was_realigned(R, 0);
  hipstr_left_aligned* H = new hipstr_left_aligned();
  auto run_jobs = [&](const std::vector<int>& jobs) -> hipstr_status_t {
    if (jobs.empty()) return HIPSTR_OK;
    std::vector<int32_t> ref_off{0}, read_off{0}, win_start;
    std::string refs, seqs;
    for (int r : jobs) {
      const Trimmed& a = reads[r];
      const int l = locus_of[r];
      const int32_t start = std::max(a.pos - kAlignWindowWidth - 1, 0);
      const int32_t stop = std::min<int32_t>(a.end_pos + kAlignWindowWidth - 1, (int32_t)chrom_len[l] - 1);
      if (stop < start) return HIPSTR_ERR_BAD_ARG;
      refs.append(chrom_seq[l] + start, stop - start + 1);
      seqs += a.bases;
      ref_off.push_back((int32_t)refs.size());
      read_off.push_back((int32_t)seqs.size());
      win_start.push_back(start);
    }
    int max_ref = 0, max_read = 0;
    for (size_t k = 0; k < jobs.size(); k++) {
      max_ref = std::max(max_ref, ref_off[k + 1] - ref_off[k]);
      max_read = std::max(max_read, read_off[k + 1] - read_off[k]);
    }
    const int32_t stride = max_ref + max_read + 2;
    std::vector<char> ops(jobs.size() * (size_t)stride);
    std::vector<int32_t> lens(jobs.size());
    std::vector<float> score(jobs.size());
    hipstr_status_t st = hipstr_nw_align_batch_host(ctx, (int32_t)jobs.size(), ref_off.data(), refs.data(), read_off.data(), seqs.data(), 0,
                                                    stride, ops.data(), lens.data(), score.data());
    if (st != HIPSTR_OK) return st;
    H->nw_alignments += (int64_t)jobs.size();
    parallel_for(jobs.size(), [&](size_t k) {
      realigned[jobs[k]] = finish_realign(reads[jobs[k]], win_start[k], refs.data() + ref_off[k], &ops[k * (size_t)stride], lens[k]);
      was_realigned[jobs[k]] = 1;
    });
    return HIPSTR_OK;
  };
  std::vector<int> jobs;
  for (int l = 0; l < n_loci; l++)
    for (const auto& kv : occurrences[l])
      if (!matches_reference(reads[kv.second[0]])) jobs.push_back(kv.second[0]);
  hipstr_status_t st = run_jobs(jobs);
  if (st != HIPSTR_OK) { delete H; return st; }
  jobs.clear();
  for (int l = 0; l < n_loci; l++)
    for (const auto& kv : occurrences[l]) {
      const int first = kv.second[0];
      if (!was_realigned[first] || !realigned[first].ok || realigned[first].bases.size() == kv.first.size()) continue;
      for (size_t k = 1; k < kv.second.size(); k++)
        if (!matches_reference(reads[kv.second[k]])) jobs.push_back(kv.second[k]);
    }
  st = run_jobs(jobs);
  if (st != HIPSTR_OK) { delete H; return st; }

  // the reference's loop, read by read (genotyper_bam_processor.cpp:50-93), one locus per task; the per-locus pieces are
  // then laid end to end
  struct Piece {
    std::vector<int32_t> seq_len, read_start, read_stop, cigar_n, cigar_len, sample_label, name_id, source;
    std::vector<char> bases, quals, cigar_type;
    std::vector<double> log_p1, log_p2;
    std::vector<uint8_t> rev_strand, use_for_haps;
    int64_t failed = 0;
  };
  std::vector<Piece> pieces(n_loci);
  parallel_for((size_t)n_loci, [&](size_t li) {
    const int l = (int)li;
    Piece& P = pieces[l];
    std::map<std::string, Aligned> seq_to_aln;   // the alignment every later read with this sequence reuses
    for (int r = raw->locus_read_off[l]; r < raw->locus_read_off[l + 1]; r++) {
      const Trimmed& a = reads[r];
      if (a.bases.empty()) continue;
      auto prev = seq_to_aln.find(a.bases);
      Aligned result;
      if (prev != seq_to_aln.end() && prev->second.bases.size() == a.bases.size()) {
        result = prev->second;       // start / stop / CIGAR of the earlier read, this read's own qualities and bases
        result.quals = a.quals;
        result.bases = a.bases;
        for (char& c : result.bases) c = up(c);
      } else {
        if (matches_reference(a)) result = convert_alignment(a, chrom_seq[l]);
        else {
          if (!was_realigned[r] || !realigned[r].ok) { P.failed++; continue; }
          result = realigned[r];
        }
        seq_to_aln[a.bases] = result;
      }
      P.bases.insert(P.bases.end(), result.bases.begin(), result.bases.end());
      P.quals.insert(P.quals.end(), result.quals.begin(), result.quals.end());
      P.seq_len.push_back((int32_t)result.bases.size());
      P.read_start.push_back(result.start);
      P.read_stop.push_back(result.stop);
      for (const auto& op : result.cigar) { P.cigar_type.push_back(op.first); P.cigar_len.push_back(op.second); }
      P.cigar_n.push_back((int32_t)result.cigar.size());
      P.sample_label.push_back(raw->sample_label[r]);
      P.name_id.push_back(raw->name_id[r]);
      P.log_p1.push_back(raw->log_p1[r]);
      P.log_p2.push_back(raw->log_p2[r]);
      P.rev_strand.push_back(raw->rev_strand ? raw->rev_strand[r] : 0);
      P.use_for_haps.push_back(raw->use_for_haps ? raw->use_for_haps[r] : 1);
      P.source.push_back(r);
    }
  });
  H->locus_read_off.push_back(0);
  H->locus_sample_off.assign(raw->locus_sample_off, raw->locus_sample_off + n_loci + 1);
  H->read_seq_off.push_back(0);
  H->cigar_off.push_back(0);
  auto append = [](auto& dst, const auto& src) { dst.insert(dst.end(), src.begin(), src.end()); };
  for (int l = 0; l < n_loci; l++) {
    const Piece& P = pieces[l];
    for (int32_t n : P.seq_len) H->read_seq_off.push_back(H->read_seq_off.back() + n);
    for (int32_t n : P.cigar_n) H->cigar_off.push_back(H->cigar_off.back() + n);
    append(H->bases, P.bases); append(H->quals, P.quals); append(H->cigar_type, P.cigar_type); append(H->cigar_len, P.cigar_len);
    append(H->read_start, P.read_start); append(H->read_stop, P.read_stop); append(H->sample_label, P.sample_label);
    append(H->name_id, P.name_id); append(H->log_p1, P.log_p1); append(H->log_p2, P.log_p2); append(H->rev_strand, P.rev_strand);
    append(H->use_for_haps, P.use_for_haps); append(H->source, P.source);
    H->fail_count += P.failed;
    H->locus_read_off.push_back((int32_t)H->read_start.size());
    H->haploid.push_back(raw->haploid ? raw->haploid[l] : 0);
  }
  hipstr_locus_reads_t& v = H->view;
  v.locus_read_off = H->locus_read_off.data();
  v.locus_sample_off = H->locus_sample_off.data();
  v.read_seq_off = H->read_seq_off.data();
  v.bases = H->bases.data();
  v.quals = H->quals.data();
  v.read_start = H->read_start.data();
  v.cigar_off = H->cigar_off.data();
  v.cigar_type = H->cigar_type.data();
  v.cigar_len = H->cigar_len.data();
  v.sample_label = H->sample_label.data();
  v.name_id = H->name_id.data();
  v.log_p1 = H->log_p1.data();
  v.log_p2 = H->log_p2.data();
  v.haploid = H->haploid.data();
  v.rev_strand = H->rev_strand.data();
  v.read_stop = H->read_stop.data();
  v.use_for_haps = H->use_for_haps.data();
  *out_handle = H;
  return HIPSTR_OK;
}

/* One read through the host steps (TrimAlignment, then convertAlignment or the tail of realign()); see the header. */
int32_t hipstr_left_align_one(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar,
                              const char* cigar_type, const int32_t* cigar_len, const char* chrom_seq, int32_t do_trim,
                              int32_t trim_start, int32_t trim_stop, const char* nw_ops, int32_t* window, int32_t* out_pos,
                              char* out_seq, char* out_qual, int32_t* n_out_cigar, char* out_ctype, int32_t* out_clen) {
  using namespace hipstr;
  if (!bases || !quals || !cigar_type || !cigar_len || !chrom_seq || !window || !out_pos || !out_seq || !out_qual || !n_out_cigar ||
      !out_ctype || !out_clen)
    return -2;
  Trimmed a;
  a.pos = pos; a.end_pos = end_pos; a.bases = bases; a.quals = quals;
  for (int c = 0; c < n_cigar; c++) a.cigar.emplace_back(cigar_type[c], cigar_len[c]);
  if (do_trim) trim_alignment(a, trim_start, trim_stop);
  if (a.bases.empty()) return -1;
  Aligned r;
  int32_t how;
  if (matches_reference(a)) { r = convert_alignment(a, chrom_seq); how = 1; }
  else {
    const int32_t chrom_len = (int32_t)std::strlen(chrom_seq);
    window[0] = std::max(a.pos - kAlignWindowWidth - 1, 0);
    window[1] = std::min(a.end_pos + kAlignWindowWidth - 1, chrom_len - 1) - window[0] + 1;
    std::memcpy(out_seq, a.bases.c_str(), a.bases.size() + 1);   // the (trimmed) read to align against the window
    if (!nw_ops) return 3;
    r = finish_realign(a, window[0], chrom_seq + window[0], nw_ops, (int32_t)std::strlen(nw_ops));
    how = r.ok ? 2 : 0;
  }
  out_pos[0] = r.start; out_pos[1] = r.stop;
  std::memcpy(out_seq, r.bases.c_str(), r.bases.size() + 1);
  std::memcpy(out_qual, r.quals.c_str(), r.quals.size() + 1);
  *n_out_cigar = (int32_t)r.cigar.size();
  for (size_t k = 0; k < r.cigar.size(); k++) { out_ctype[k] = r.cigar[k].first; out_clen[k] = r.cigar[k].second; }
  return how;
}

const hipstr_locus_reads_t* hipstr_left_aligned_reads(const hipstr_left_aligned_t* h) { return h ? &h->view : nullptr; }
const int32_t* hipstr_left_aligned_source(const hipstr_left_aligned_t* h, int64_t* n_reads) {
  if (!h) return nullptr;
  if (n_reads) *n_reads = (int64_t)h->source.size();
  return h->source.data();
}
void hipstr_left_aligned_counts(const hipstr_left_aligned_t* h, int64_t* failed, int64_t* nw_alignments) {
  if (!h) return;
  if (failed) *failed = h->fail_count;
  if (nw_alignments) *nw_alignments = h->nw_alignments;
}
void hipstr_left_aligned_free(hipstr_left_aligned_t* h) { delete h; }

}  // extern "C"
