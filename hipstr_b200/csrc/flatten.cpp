/*
 * flatten.cpp -- lowers a batch of loci to the device layout (layout.h).
 *
 * What the reference does lazily inside its object graph is done here once per batch:
 *   - haplotype enumeration (Haplotype.cpp:123-206) -> explicit oriented sequences;
 *   - homopolymer classes of every haplotype row (Haplotype.cpp:239-287, HapBlock.cpp:7-30,
 *     HapAligner.cpp:119-120), INCLUDING the history the reference's DP-row reuse leaves in them
 *     (HapAligner.cpp:54-60, 612-634): rows of blocks left of the last changed
 *     block keep the classes they got under an earlier haplotype of the same aligned run;
 *   - StutterAlignerClass constructor tables (StutterAlignerClass.h:35-80) and the 13
 *     log_prob_pcr_artifact values per repeat allele (RepeatStutterInfo.h:53-61,
 *     stutter_model.cpp:29-53), with glibc log().
 * With those classes baked in, an independent from-scratch DP per (read, haplotype) is
 * bit-identical to the reference's incremental one.
 */
#include "flatten.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace hipstr {

HostTables::HostTables() {
  int_logs[0] = -1000;
  for (int i = 1; i < 10000; i++) int_logs[i] = std::log(i);
  double ok[42], bad[42];
  ok[0] = -100000;
  bad[0] = -std::log(3);
  for (int q = 1; q <= 41; q++) {
    ok[q] = std::log(1.0 - std::pow(10.0, q / (-10.0)));
    bad[q] = std::log(std::pow(10.0, q / (-10.0)) / 3.0);
  }
  for (int byte = 0; byte < 256; byte++) {
    const int c = (signed char)byte;             // the reference compares plain (signed) chars
    const int q = c < '!' ? 0 : (c > 'J' ? 41 : c - '!');
    qual_lut[byte][0] = ok[q];
    qual_lut[byte][1] = bad[q];
  }
  const double dindel[10] = {2.9e-5, 2.9e-5, 2.9e-5, 2.9e-5, 4.3e-5, 1.1e-4, 2.4e-4, 5.7e-4, 1.0e-3, 1.4e-3};
  for (int k = 0; k < 3; k++) trans[k][0] = 0.0;
  for (int h = 1; h <= 15; h++) {
    const double gap = h <= 10 ? std::log(dindel[h - 1]) : std::log(dindel[9] + 4.3e-4 * (h - 10));
    trans[1][h] = gap;
    trans[2][h] = gap;
    trans[0][h] = std::log(1.0 - std::exp(trans[1][h]) - std::exp(trans[2][h]));
  }
  log_one_half = std::log(0.5);
}

const HostTables& host_tables() {
  static const HostTables t;
  return t;
}

void haplotype_options(int n_blocks, const int32_t* n_opts, int64_t hap, int32_t* out) {
  int64_t stride = 1;
  for (int b = 0; b < n_blocks; b++) {
    const int64_t q = hap / stride;
    const int digit = (int)(q % n_opts[b]);
    const bool reflected = ((q / n_opts[b]) & 1) != 0;
    out[b] = reflected ? n_opts[b] - 1 - digit : digit;
    stride *= n_opts[b];
  }
}

namespace {

// Within-block homopolymer run lengths of one oriented option.  The reference fills the left
// runs then the right runs with a single counter it never resets (HapBlock.cpp:18-28), so the
// right runs inherit the left-run length of the last base; kept because it is observable.
struct Runs {
  std::vector<int> left, right;
  void build(const std::string& s) {
    const int n = (int)s.size();
    left.assign(n, 0);
    right.assign(n, 0);
    int c = 0;
    for (int j = 1; j < n; j++) left[j] = c = (s[j - 1] == s[j] ? c + 1 : 0);
    for (int j = n - 2; j >= 0; j--) right[j] = c = (s[j + 1] == s[j] ? c + 1 : 0);
  }
};

struct Option {
  std::string seq[2];   // forward, reversed
  Runs runs[2];
  int rep[2] = {-1, -1};
};

struct BlockInfo {
  int period;
  std::vector<Option> opts;
};

// Homopolymer length around base `pos` of oriented block `b` given the current choice of options
// (Haplotype::homopolymer_length, Haplotype.cpp:277-287).  `blk[k]` / `opt[k]` are in the
// orientation's own block order.
int homopolymer(const std::vector<const Option*>& opt, int side, int b, int pos) {
  const std::string& s = opt[b]->seq[side];
  const Runs& r = opt[b]->runs[side];
  int l = r.left[pos], rt = r.right[pos];
  const char c = s[pos];
  if (pos - l == 0) {
    for (int k = b - 1; k >= 0; k--) {
      const std::string& t = opt[k]->seq[side];
      if (t.empty()) continue;
      if (t.back() != c) break;
      const int run = opt[k]->runs[side].left[t.size() - 1];
      l += 1 + run;
      if (run != (int)t.size()) break;
    }
  }
  if (pos + rt == (int)s.size() - 1) {
    for (int k = b + 1; k < (int)opt.size(); k++) {
      const std::string& t = opt[k]->seq[side];
      if (t.empty()) continue;
      if (t[0] != c) break;
      const int run = opt[k]->runs[side].right[0];
      rt += 1 + run;
      if (run != (int)t.size()) break;
    }
  }
  return l + rt + 1;
}

int pick_variant(int n_left, int n_right) {
  for (int v = 0; v < kNumColVariants; v++) {
    const int c = kColVariants[v];
    if ((n_left + c - 1) / c + (n_right + c - 1) / c <= 32) return v;
  }
  return -1;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Bases travel to the device as codes 0..4 = A,C,G,T,N.  The reference compares raw characters
// (HapAligner.cpp:115,149); restricted to this alphabet that is the same relation.
inline int base_code(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    case 'N': return 4;
    default: return -1;
  }
}
// appends the codes of [b, e) to dst; false if a character is outside ACGTN
template <class It, class V>
bool append_codes(V& dst, It b, It e) {
  for (; b != e; ++b) {
    const int x = base_code(*b);
    if (x < 0) return false;
    dst.push_back((typename V::value_type)x);
  }
  return true;
}

}  // namespace

int64_t count_alignments(const hipstr_align_batch_t* b) {
  int64_t total = 0;
  for (int l = 0; l < b->n_loci; l++) {
    int64_t H = 1;
    for (int k = b->locus_block_off[l]; k < b->locus_block_off[l + 1]; k++)
      H *= b->block_opt_off[k + 1] - b->block_opt_off[k];
    int64_t live = H;
    if (b->realign_hap) {
      live = 0;
      for (int64_t h = 0; h < H; h++) live += b->realign_hap[b->locus_hap_off[l] + h] != 0;
    }
    for (int p = b->locus_pool_off[l]; p < b->locus_pool_off[l + 1]; p++)
      if ((!b->realign_pool || b->realign_pool[p]) && b->pool_seed[p] >= 0) total += live;
  }
  return total;
}

hipstr_status_t flatten_batch(const hipstr_align_batch_t* b, FlatBatch& out, std::string& err) {
  if (!b || b->n_loci < 0) { err = "null batch"; return HIPSTR_ERR_BAD_ARG; }
  if (b->n_loci > 0 && (!b->locus_block_off || !b->locus_pool_off || !b->locus_hap_off || !b->locus_out_off ||
                        !b->block_period || !b->block_opt_off || !b->block_stutter || !b->opt_seq_off || !b->opt_seq ||
                        !b->pool_seq_off || !b->pool_bases || !b->pool_quals || !b->pool_seed)) {
    err = "null array in batch";
    return HIPSTR_ERR_BAD_ARG;
  }
  for (int v = 0; v < kNumColVariants; v++) { out.n_max[v] = 16; out.l_max[v] = 2; }
  out.n_out = b->n_loci ? b->locus_out_off[b->n_loci] : 0;

  int64_t total_pairs = count_alignments(b);
  out.n_alignments = total_pairs;

  for (int l = 0; l < b->n_loci; l++) {
    const int b0 = b->locus_block_off[l], nb = b->locus_block_off[l + 1] - b0;
    if (nb < 1 || nb > HIPSTR_MAX_BLOCKS) { err = "locus needs 1.." + std::to_string(HIPSTR_MAX_BLOCKS) + " blocks"; return HIPSTR_ERR_UNSUPPORTED; }
    if (b->block_period[b0] != 0 || b->block_period[b0 + nb - 1] != 0) {
      err = "first and last haplotype block must be flank blocks";   // compute_aln_logprob assumes it
      return HIPSTR_ERR_UNSUPPORTED;
    }
    std::vector<BlockInfo> blk(nb);
    std::vector<int32_t> n_opts(nb);
    int64_t H = 1;
    for (int k = 0; k < nb; k++) {
      const int o0 = b->block_opt_off[b0 + k], o1 = b->block_opt_off[b0 + k + 1];
      if (o1 <= o0) { err = "block without options"; return HIPSTR_ERR_BAD_ARG; }
      blk[k].period = b->block_period[b0 + k];
      if (blk[k].period < 0 || blk[k].period > 9) { err = "motif period must be 1..9"; return HIPSTR_ERR_BAD_ARG; }
      n_opts[k] = o1 - o0;
      H *= n_opts[k];
      if (H > (1 << 20)) { err = "too many haplotypes"; return HIPSTR_ERR_UNSUPPORTED; }
      blk[k].opts.resize(o1 - o0);
      for (int o = o0; o < o1; o++) {
        Option& op = blk[k].opts[o - o0];
        op.seq[0].assign(b->opt_seq + b->opt_seq_off[o], b->opt_seq + b->opt_seq_off[o + 1]);
        if (op.seq[0].empty()) { err = "empty block option"; return HIPSTR_ERR_UNSUPPORTED; }
        if (op.seq[0].size() > 60000) { err = "block option too long"; return HIPSTR_ERR_UNSUPPORTED; }
        op.seq[1].assign(op.seq[0].rbegin(), op.seq[0].rend());
        for (int s = 0; s < 2; s++) op.runs[s].build(op.seq[s]);
        if (blk[k].period > 0) {
          // StutterAlignerClass constructor (StutterAlignerClass.h:49-80) for both orientations
          const int p = blk[k].period, B = (int)op.seq[0].size();
          int n_del = HIPSTR_MAX_ARTIFACT_UNITS;
          while (n_del * p > B) n_del--;
          const double* prm = b->block_stutter + 6 * (size_t)(b0 + k);
          const double in_step = std::log(1 - prm[0]), in_nostep = std::log(prm[0]);
          const double in_up = std::log(prm[1]), in_down = std::log(prm[2]);
          const double equal = std::log(1 - prm[1] - prm[2] - prm[4] - prm[5]);
          for (int s = 0; s < 2; s++) {
            DevRep r;
            std::memset(&r, 0, sizeof(r));
            r.seq_off = (int32_t)out.hapbytes.size();
            if (!append_codes(out.hapbytes, op.seq[s].begin(), op.seq[s].end())) { err = "haplotype bases must be A,C,G,T or N"; return HIPSTR_ERR_UNSUPPORTED; }
            r.len = B;
            r.period = p;
            r.n_del = n_del;
            r.left_align = (s == 0);
            r.runs_off = (int32_t)out.runs.size();
            const int n_lag = std::max(n_del, 1);
            for (int k2 = 1; k2 <= n_lag; k2++) {
              const int lag = k2 * p;
              size_t base = out.runs.size();
              out.runs.resize(base + B, 0);
              for (int i = lag; i < B; i++)
                out.runs[base + i] = op.seq[s][i - lag] != op.seq[s][i] ? 0 : (uint16_t)(1 + out.runs[base + i - 1]);
            }
            for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) {
              const int units = a - HIPSTR_MAX_ARTIFACT_UNITS;
              double v;
              if (units == 0) v = equal;
              else if (units > 0) v = in_up + in_nostep + in_step * (units - 1);
              else v = (B + units * p < 0) ? -10e6 : in_down + in_nostep + in_step * (-units - 1);
              r.art[a] = v;
            }
            op.rep[s] = (int)out.reps.size();
            out.reps.push_back(r);
          }
        }
      }
    }
    if (b->locus_hap_off[l + 1] - b->locus_hap_off[l] != H) { err = "locus_hap_off inconsistent with block options"; return HIPSTR_ERR_BAD_ARG; }
    const int P = b->locus_pool_off[l + 1] - b->locus_pool_off[l];
    if (b->locus_out_off[l + 1] - b->locus_out_off[l] != (int64_t)P * H) { err = "locus_out_off inconsistent"; return HIPSTR_ERR_BAD_ARG; }
    const uint8_t* mask = b->realign_hap ? b->realign_hap + b->locus_hap_off[l] : nullptr;
    if (mask) out.hap_mask.insert(out.hap_mask.end(), mask, mask + H);   // dense, parallel to hapsides / 2

    // ---- haplotypes: oriented sequences + row classes with the reference's reuse history ----
    const int hap_rec0 = (int)out.hapsides.size();
    std::vector<int32_t> cur(nb), prev(nb);
    std::vector<std::vector<uint8_t> > cached[2];   // [side][oriented block] -> classes of its rows
    cached[0].resize(nb); cached[1].resize(nb);
    std::map<std::string, int> seg1_ids[2];
    bool reuse = false;
    int max_len = 0;
    for (int64_t h = 0; h < H; h++) {
      haplotype_options(nb, n_opts.data(), h, cur.data());
      int last_changed = -1;
      if (h > 0) for (int k = 0; k < nb; k++) if (cur[k] != prev[k]) last_changed = k;
      prev = cur;
      const bool live = !mask || mask[h];
      for (int side = 0; side < 2; side++) {
        DevHapSide hs;
        std::memset(&hs, 0, sizeof(hs));
        if (!live) { out.hapsides.push_back(hs); continue; }   // placeholder keeps indexing dense
        std::vector<const Option*> opt(nb);
        std::vector<int> period(nb);
        for (int k = 0; k < nb; k++) {
          const int src = side == 0 ? k : nb - 1 - k;
          opt[k] = &blk[src].opts[cur[src]];
          period[k] = blk[src].period;
        }
        const int changed = last_changed < 0 ? -1 : (side == 0 ? last_changed : nb - 1 - last_changed);
        hs.seq_off = (int32_t)out.hapbytes.size();
        int len = 0;
        for (int k = 0; k < nb; k++) {
          if (!append_codes(out.hapbytes, opt[k]->seq[side].begin(), opt[k]->seq[side].end())) { err = "haplotype bases must be A,C,G,T or N"; return HIPSTR_ERR_UNSUPPORTED; }
          len += (int)opt[k]->seq[side].size();
        }
        hs.len = len;
        hs.row_off = (int32_t)out.hapbytes.size();
        out.hapbytes.resize(out.hapbytes.size() + len, 0);
        uint8_t* rows = out.hapbytes.data() + hs.row_off;
        hs.blk_off = (int32_t)out.blocks.size();
        hs.n_blocks = nb;
        int row = 0, first_repeat_row = -1;
        for (int k = 0; k < nb; k++) {
          const int n = (int)opt[k]->seq[side].size();
          DevBlock db = {row, n, period[k] > 0 ? opt[k]->rep[side] : -1, 0};
          out.blocks.push_back(db);
          if (period[k] > 0) {
            if (first_repeat_row < 0) first_repeat_row = row;
            for (int c = 0; c < n; c++) rows[row + c] = HIPSTR_ROW_REPEAT;
          } else {
            hs.n_seed_pos += n;
            std::vector<uint8_t>& keep = cached[side][k];
            if (!(reuse && k < changed)) {   // HapAligner.cpp:54-60: these rows are recomputed now
              keep.resize(n);
              for (int c = 0; c < n; c++) {
                const int hp = std::max(homopolymer(opt, side, k, c), homopolymer(opt, side, k, std::max(0, c - 1)));
                keep[c] = (uint8_t)std::min(15, hp);
              }
            }
            for (int c = 0; c < n; c++) rows[row + c] = keep[c];
            if (k > 0 && period[k - 1] > 0) rows[row] |= HIPSTR_ROW_AFTER_REPEAT;
          }
          row += n;
        }
        hs.first_rep = nb;
        for (int k = nb - 1; k >= 0; k--) if (period[k] > 0) hs.first_rep = k;
        hs.seg1_class = -1;
        if (nb == 3 && period[0] == 0 && period[1] > 0 && period[2] == 0) {   // the canonical HipSTR haplotype
          std::string key((const char*)out.hapbytes.data() + hs.seq_off, first_repeat_row);
          key.append((const char*)rows, first_repeat_row);
          auto it = seg1_ids[side].find(key);
          if (it == seg1_ids[side].end()) it = seg1_ids[side].emplace(key, (int)seg1_ids[side].size()).first;
          hs.seg1_class = it->second;
        }
        max_len = std::max(max_len, len);
        out.hapsides.push_back(hs);
      }
      reuse = live;   // HapAligner.cpp:615-619,634: a skipped haplotype breaks the reuse chain
    }
    // DevHapSide offsets into hapbytes were taken while the vector was still growing: they are
    // indices, not pointers, so they stay valid.

    // ---- pooled reads and jobs ----
    int64_t live_haps = H;
    if (mask) { live_haps = 0; for (int64_t h = 0; h < H; h++) live_haps += mask[h] != 0; }
    // Split a pool's haplotypes over several warps only when the batch is too small to fill the GPU.
    int chunk = (int)H;
    if (total_pairs > 0 && total_pairs / H < 16384) chunk = (int)std::max<int64_t>(1, std::min<int64_t>(H, total_pairs / 16384));
    for (int p = b->locus_pool_off[l]; p < b->locus_pool_off[l + 1]; p++) {
      const int s0 = b->pool_seq_off[p], n = b->pool_seq_off[p + 1] - s0;
      if (n < 0) { err = "pool_seq_off not monotone"; return HIPSTR_ERR_BAD_ARG; }
      DevPool dp;
      dp.seq_off = (int32_t)out.bases.size();
      dp.len = n;
      dp.seed = b->pool_seed[p];
      dp.locus = l;
      dp.out_off = b->locus_out_off[l] + (int64_t)(p - b->locus_pool_off[l]) * H;
      dp.hap_rec0 = hap_rec0;
      dp.n_haps = (int32_t)H;
      const int padded = round_up(std::max(n, 1), 16);
      if (!append_codes(out.bases, b->pool_bases + s0, b->pool_bases + s0 + n)) { err = "read bases must be A,C,G,T or N"; return HIPSTR_ERR_UNSUPPORTED; }
      out.bases.resize(out.bases.size() + (padded - n), 4);
      out.quals.insert(out.quals.end(), b->pool_quals + s0, b->pool_quals + s0 + n);
      out.quals.resize(out.quals.size() + (padded - n), '!');
      const int pool_id = (int)out.pools.size();
      out.pools.push_back(dp);
      if (b->realign_pool && !b->realign_pool[p]) continue;
      if (dp.seed < 0) {   // HapAligner.cpp:333-337: LL 0 for every haplotype, mask ignored
        DevJob j = {pool_id, 0, (int32_t)H, 0};
        out.jobs[0].push_back(j);
        continue;
      }
      if (dp.seed == 0 || dp.seed >= n - 1) { err = "invalid alignment seed"; return HIPSTR_ERR_INVALID_SEED; }
      if (live_haps == 0) continue;
      const int v = pick_variant(dp.seed, n - dp.seed - 1);
      if (v < 0) { err = "read longer than the kernel's limit"; return HIPSTR_ERR_UNSUPPORTED; }
      out.n_max[v] = std::max(out.n_max[v], padded);
      out.l_max[v] = std::max(out.l_max[v], round_up(max_len, 2));
      for (int h0 = 0; h0 < H; h0 += chunk) {
        DevJob j = {pool_id, h0, (int32_t)std::min<int64_t>(H, h0 + chunk), 0};
        out.jobs[v].push_back(j);
      }
    }
  }
  return HIPSTR_OK;
}

}  // namespace hipstr
