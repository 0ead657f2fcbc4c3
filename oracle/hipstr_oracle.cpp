/*
 * hipstr_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see hipstr_oracle.h).
 *
 * Serial CPU restatement of HipSTR's read-vs-haplotype HMM and genotype
 * posterior reduction, written against flat arrays.  Citations are
 * path:line in the reference checkout (tfwillems/HipSTR @ b2033bf).
 *
 * Build: g++ -O2 -ffp-contract=off (no -march=native / -ffast-math: the
 * single-precision bit tricks below must round exactly like the reference's
 * x86-64 SSE2 build).
 */
#include "hipstr_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Approximate single-precision exp/log (third-party fastapprox, vendored by the
// reference as src/fastonebigheader.h).  Every operation is a separately
// rounded binary32 op.
// ---------------------------------------------------------------------------
inline uint32_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float float_of(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// fasterpow2 / fasterexp (fastonebigheader.h:206-218)
inline float coarse_exp(float p) {
  float x = 1.442695040f * p;
  float c = (x < -126) ? -126.0f : x;
  return float_of(static_cast<uint32_t>(8388608.0f * (c + 126.94269504f)));
}
// fasterlog (fastonebigheader.h:348-357)
inline float coarse_log(float x) {
  float y = static_cast<float>(bits_of(x));
  y *= 8.2629582881927490e-8f;
  return y - 87.989971088f;
}
// fastpow2 / fastexp (fastonebigheader.h:188-204)
inline float fine_exp(float p) {
  float x = 1.442695040f * p;
  float offset = (x < 0) ? 1.0f : 0.0f;
  float c = (x < -126) ? -126.0f : x;
  int w = static_cast<int>(c);
  float z = c - w + offset;
  return float_of(static_cast<uint32_t>(
      8388608.0f * (c + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z)));
}
// fastlog2 / fastlog (fastonebigheader.h:320-337)
inline float fine_log(float x) {
  uint32_t xi = bits_of(x);
  float mx = float_of((xi & 0x007FFFFFu) | 0x3f000000u);
  float y = static_cast<float>(xi);
  y *= 1.1920928955078125e-7f;
  float l2 = y - 124.22551499f - 1.498030302f * mx - 1.72587999f / (0.3520887068f + mx);
  return 0.69314718f * l2;
}

// ---------------------------------------------------------------------------
// Constant tables, all from glibc libm exactly as the reference builds them.
// ---------------------------------------------------------------------------
const double kImpossible = -1000000000;      // HapAligner.cpp:20
const double kLargeNegative = -10e6;         // RepeatStutterInfo.h:12
const double kInsToIns = -1.0, kDelToDel = -1.0;                       // AlignmentModel.h:7,9
const double kInsToMatch = -0.4586751453870818910216436;               // AlignmentModel.h:8
const double kDelToMatch = -0.4586751453870818910216436;               // AlignmentModel.h:10
const int kMaxHomop = 15;                                               // AlignmentModel.h:6
const int kMinSeedDist = 5;                                             // HapAligner.cpp:17

struct Tables {
  double int_logs[10000];
  double qual_correct[256], qual_error[256];  // indexed by (unsigned char) quality
  double m2m[16], m2i[16], m2d[16];
  double log_thresh, log_half;
  Tables() {
    int_logs[0] = -1000;                                   // mathops.cpp:16
    for (int i = 1; i < 10000; i++) int_logs[i] = std::log(i);
    // base_quality.h:29-38, clamping :45-75 ('!'..'J'; signed char compare)
    double corr[42], err[42];
    corr[0] = -100000; err[0] = -std::log(3);
    for (int i = 1; i <= 41; i++) {
      corr[i] = std::log(1.0 - std::pow(10.0, i / (-10.0)));
      err[i] = std::log(std::pow(10.0, i / (-10.0)) / 3.0);
    }
    for (int c = 0; c < 256; c++) {
      int sc = static_cast<signed char>(static_cast<unsigned char>(c));
      int idx = sc < '!' ? 0 : (sc > 'J' ? 41 : sc - '!');
      qual_correct[c] = corr[idx];
      qual_error[c] = err[idx];
    }
    // AlignmentModel.cpp:9,20-32
    const double dindel[10] = {2.9e-5, 2.9e-5, 2.9e-5, 2.9e-5, 4.3e-5, 1.1e-4, 2.4e-4, 5.7e-4, 1.0e-3, 1.4e-3};
    m2m[0] = m2i[0] = m2d[0] = 0;
    for (int i = 1; i <= kMaxHomop; i++) {
      m2i[i] = (i <= 10 ? std::log(dindel[i - 1]) : std::log(dindel[9] + (4.3e-4) * (i - 10)));
      m2d[i] = m2i[i];
      m2m[i] = std::log(1.0 - std::exp(m2i[i]) - std::exp(m2d[i]));
    }
    log_thresh = std::log(0.001);   // mathops.h:36
    log_half = std::log(0.5);       // mathops.cpp:9
  }
};
const Tables& T() { static Tables t; return t; }

// mathops.cpp:97-106
double lse_vec(const double* v, int n) {
  double mx = v[0];
  for (int i = 1; i < n; i++) if (v[i] > mx) mx = v[i];     // std::max_element: first max
  double total = 0;
  for (int i = 0; i < n; i++) {
    double diff = v[i] - mx;
    if (diff > T().log_thresh) total += coarse_exp(static_cast<float>(diff));
  }
  return mx + coarse_log(static_cast<float>(total));
}
// mathops.cpp:86-95
double lse2(double a, double b) {
  double hi = a > b ? a : b, lo = a > b ? b : a;
  double diff = lo - hi;
  return diff < T().log_thresh ? hi : hi + fine_log(1 + fine_exp(static_cast<float>(diff)));
}
// mathops.cpp:44-49 (exact)
double exact_lse(const double* b, const double* e) {
  double mx = *std::max_element(b, e), total = 0.0;
  for (const double* p = b; p != e; ++p) total += std::exp(*p - mx);
  return mx + std::log(total);
}

// stutter_model.h:36-63 + stutter_model.cpp:29-53
struct StutterPmf {
  double in_nostep, in_step, in_up, in_down, out_nostep, out_step, out_up, out_down, equal;
  int period;
  StutterPmf(const double* prm, int p) : period(p) {
    in_step = std::log(1 - prm[0]); in_nostep = std::log(prm[0]);
    in_up = std::log(prm[1]); in_down = std::log(prm[2]);
    out_step = std::log(1 - prm[3]); out_nostep = std::log(prm[3]);
    out_up = std::log(prm[4]); out_down = std::log(prm[5]);
    equal = std::log(1 - prm[1] - prm[2] - prm[4] - prm[5]);
  }
  double operator()(int sample_bps, int read_bps) const {
    int d = read_bps - sample_bps;
    if (d % period != 0) {
      int eff = d - d / period;
      return eff < 0 ? out_down + out_nostep + out_step * (-eff - 1) : out_up + out_nostep + out_step * (eff - 1);
    }
    int reps = d / period;
    if (reps == 0) return equal;
    return reps < 0 ? in_down + in_nostep + in_step * (-reps - 1) : in_up + in_nostep + in_step * (reps - 1);
  }
};

// ---------------------------------------------------------------------------
// Haplotype structure of one locus, one orientation.
// ---------------------------------------------------------------------------
struct OrientedBlock {
  std::string seq;   // current option, already reversed for the reverse orientation
  int period;        // 0 = flank
  int opt;           // option index in the block
  int src_block;     // index in the batch's block arrays
  std::vector<int> lrun, rrun;  // within-block homopolymer run lengths (HapBlock.cpp:7-30)
};

// HapBlock::calc_homopolymer_lengths (HapBlock.cpp:7-30).  The reference uses ONE counter for
// both sweeps and does not reset it in between, so the right-run lengths of a block that ends in
// a homopolymer run start from the left-run length of its last base.  That is part of the
// reference's results (it feeds the homopolymer class of the transition tables) and is kept.
void run_lengths(OrientedBlock& b) {
  size_t n = b.seq.size();
  b.lrun.assign(n, 0); b.rrun.assign(n, 0);
  if (n == 0) return;
  int count = 0;
  for (size_t j = 1; j < n; j++) { count = (b.seq[j - 1] == b.seq[j]) ? count + 1 : 0; b.lrun[j] = count; }
  for (size_t j = n - 1; j-- > 0;) { count = (b.seq[j + 1] == b.seq[j]) ? count + 1 : 0; b.rrun[j] = count; }
}

// Haplotype::left_homopolymer_len / right_homopolymer_len (Haplotype.cpp:239-275): walk to the
// nearest non-empty neighbour block; it contributes 1 + its boundary run if its boundary base is c.
// The walk continues past a block only when that run length EQUALS the block size, which a left
// run never does and a right run does only through the carried counter above (a 2-base block "XX").
int neighbour_run(const std::vector<OrientedBlock>& hb, int b, int step, char c) {
  int total = 0;
  for (; b >= 0 && b < (int)hb.size(); b += step) {
    const std::string& s = hb[b].seq;
    if (s.empty()) continue;
    if (step < 0) {
      if (s.back() != c) break;
      int run = hb[b].lrun[s.size() - 1];
      total += 1 + run;
      if (run != (int)s.size()) break;
    } else {
      if (s[0] != c) break;
      int run = hb[b].rrun[0];
      total += 1 + run;
      if (run != (int)s.size()) break;
    }
  }
  return total;
}
// Haplotype::homopolymer_length (Haplotype.cpp:277-287)
int homopolymer_len(const std::vector<OrientedBlock>& hb, int b, int pos) {
  const OrientedBlock& blk = hb[b];
  int l = blk.lrun[pos], r = blk.rrun[pos];
  if (pos - l == 0) l += neighbour_run(hb, b - 1, -1, blk.seq[pos]);
  if (pos + r == (int)blk.seq.size() - 1) r += neighbour_run(hb, b + 1, +1, blk.seq[pos]);
  return l + r + 1;
}

// ---------------------------------------------------------------------------
// Repeat-block evaluator for one (read side, allele) -- StutterAlignerClass.
// Columns q = 0..n-1 are read positions of the side; s is the allele sequence
// in this orientation; everything is anchored at the right ends.
// ---------------------------------------------------------------------------
struct RepeatEval {
  const char* rd; const double* lc; const double* lw; int n;
  std::string s; int B, p, n_del;
  bool left_align;
  std::vector<std::vector<int> > lag_runs;       // upstream_match_lengths_ (StutterAlignerClass.h:35-42,70-75)
  std::vector<double> match, del, ins;           // load_read tables (StutterAlignerClass.cpp:12-53)
  std::vector<double> terms;

  double emit(int q, int b) const { return rd[q] == s[b] ? lc[q] : lw[q]; }

  RepeatEval(const std::string& seq, int period, bool left_al) : s(seq), B((int)seq.size()), p(period), left_align(left_al) {
    n_del = HIPSTR_MAX_ARTIFACT_UNITS;
    while (n_del * p > B) n_del--;
    for (int k = 1; k <= n_del; k++) lag_runs.push_back(runs_for_lag(k * p));
    if (n_del == 0) lag_runs.push_back(B == 0 ? std::vector<int>() : runs_for_lag(p));
  }
  std::vector<int> runs_for_lag(int lag) const {
    std::vector<int> m(B, 0);
    for (int i = lag; i < B; i++) m[i] = (s[i - lag] != s[i]) ? 0 : 1 + m[i - 1];
    return m;
  }
  void load(const char* read, const double* lcp, const double* lwp, int len) {
    rd = read; lc = lcp; lw = lwp; n = len;
    match.assign(n, 0.0); del.assign((size_t)n * std::max(n_del, 1), 0.0); ins.assign((size_t)n * 6, 0.0);
    for (int q = 0; q < n; q++) {
      int avail = q + 1;
      double acc = 0.0;
      int t = 0;
      for (; t < std::min(avail, n_del * p); t++) {
        acc += emit(q - t, B - 1 - t);
        if ((t + 1) % p == 0) del[(size_t)q * n_del + (t + 1) / p - 1] = acc;
      }
      if (t < n_del * p) t = n_del * p;   // unreachable deletion slots stay unset
      for (; t < std::min(avail, B); t++) acc += emit(q - t, B - 1 - t);
      match[q] = acc;
      double acc_ins = 0.0;
      int max_ins = 6 * p;
      for (t = 0; t < std::min(max_ins, avail); t++) {
        if (t % p < B) acc_ins += emit(q - t, B - 1 - (t % p));
        else acc_ins += lc[q - t];
        if ((t + 1) % p == 0) ins[(size_t)q * 6 + (t + 1) / p - 1] = acc_ins;
      }
      for (; t < max_ins; t++)
        if ((t + 1) % p == 0) ins[(size_t)q * 6 + (t + 1) / p - 1] = acc_ins;
    }
  }
  // StutterAlignerClass.cpp:59-104
  double insertion(int base_len, int j, int D, int& best_pos) {
    terms.clear();
    const int* runs = lag_runs[0].data();
    double lp = -T().int_logs[B + 1] + ins[(size_t)j * 6 + D / p - 1] + (base_len > D ? match[j - D] : 0);
    best_pos = 0;
    double best = lp;
    terms.push_back(lp);
    int i = 0;
    for (; i > -std::min(std::max(0, base_len - D), B); i--) {
      int b = B - 1 + i;
      if (-i + p < B) {
        if (runs[b] == 0) {
          for (int idx = i - p; idx >= i - D; idx -= p) {
            lp -= emit(j + idx, b);
            lp += emit(j + idx, b - p);
          }
          terms.push_back(lp);
        } else {
          terms.push_back(T().int_logs[runs[b]] + lp);
          i -= (runs[b] - 1);
        }
      } else
        terms.push_back(lp);
      if (lp > best || (left_align && lp == best)) { best_pos = 1 - i; best = lp; }
    }
    if (i > -B) terms.push_back(T().int_logs[B + i] + lp);
    return lse_vec(terms.data(), (int)terms.size());
  }
  // StutterAlignerClass.cpp:106-150
  double deletion(int base_len, int j, int D, int& best_pos) {
    terms.clear();
    int k = -D / p;
    const int* runs = lag_runs[k - 1].data();
    double lp = -T().int_logs[B + D + 1];
    if (j - D <= n - 1)
      lp += match[j - D] - del[(size_t)(j - D) * n_del + k - 1];
    else
      for (int t = 0; t < base_len; t++) lp += emit(j - t, B - 1 - t + D);
    best_pos = 0;
    double best = lp;
    terms.push_back(lp);
    int i;
    for (i = 0; i > -base_len; i--) {
      int b = B - 1 + i;
      if (runs[b] == 0) {
        lp -= emit(j + i, b + D);
        lp += emit(j + i, b);
        terms.push_back(lp);
      } else {
        terms.push_back(T().int_logs[runs[b]] + lp);
        i -= (runs[b] - 1);
      }
      if (lp > best || (left_align && lp == best)) { best_pos = 1 - i; best = lp; }
    }
    if (-i < B + D) terms.push_back(T().int_logs[B + D + i] + lp);
    return lse_vec(terms.data(), (int)terms.size());
  }
  // StutterAlignerClass.cpp:152-162
  double eval(int base_len, int j, int D, int& best_pos) {
    best_pos = -1;
    if (D == 0) return match[j];
    return D > 0 ? insertion(base_len, j, D, best_pos) : deletion(base_len, j, D, best_pos);
  }
};

// ---------------------------------------------------------------------------
// One locus of the batch.
// ---------------------------------------------------------------------------
struct Locus {
  const hipstr_align_batch_t* bt;
  int first_block, nb;
  std::vector<int> nopt;
  int64_t n_haps;
  int max_rows;
  Locus(const hipstr_align_batch_t* batch, int l) : bt(batch) {
    first_block = bt->locus_block_off[l];
    nb = bt->locus_block_off[l + 1] - first_block;
    n_haps = 1; max_rows = 0;
    for (int b = 0; b < nb; b++) {
      int gb = first_block + b;
      int o0 = bt->block_opt_off[gb], o1 = bt->block_opt_off[gb + 1];
      nopt.push_back(o1 - o0);
      n_haps *= (o1 - o0);
      int mx = 0;
      for (int o = o0; o < o1; o++) mx = std::max(mx, bt->opt_seq_off[o + 1] - bt->opt_seq_off[o]);
      max_rows += mx;
    }
  }
  std::string option_seq(int b, int opt) const {
    int o = bt->block_opt_off[first_block + b] + opt;
    return std::string(bt->opt_seq + bt->opt_seq_off[o], bt->opt_seq + bt->opt_seq_off[o + 1]);
  }
  int option_len(int b, int opt) const {
    int o = bt->block_opt_off[first_block + b] + opt;
    return bt->opt_seq_off[o + 1] - bt->opt_seq_off[o];
  }
};

void hap_options(int nb, const int* nopt, int64_t hap, int* out) {
  // Closed form of the reflected mixed-radix Gray code that Haplotype::next()
  // (Haplotype.cpp:157-196) walks one step at a time, block 0 fastest.
  int64_t f = 1;
  for (int b = 0; b < nb; b++) {
    int64_t q = hap / f;
    int d = (int)(q % nopt[b]);
    out[b] = ((q / nopt[b]) & 1) ? nopt[b] - 1 - d : d;
    f *= nopt[b];
  }
}

// One orientation's DP state for one read (the reference keeps full matrices
// per read and reuses rows of unchanged leading blocks, HapAligner.cpp:54-60).
struct SideDP {
  int n = 0;
  std::string rd;
  std::vector<double> lc, lw;
  std::vector<double> M, I, D;
  std::vector<int> art_size, art_pos;
  double edge = 0.0;
};

// align_seq_to_hap (HapAligner.cpp:26-161)
void fill_side(SideDP& sd, const std::vector<OrientedBlock>& hb, const Locus& loc, bool reuse, int last_changed,
               bool reversed) {
  const Tables& t = T();
  const int n = sd.n;
  const char first = hb[0].seq[0];
  double run = 0.0;
  for (int j = 0; j < n; j++) {
    sd.M[j] = (sd.rd[j] == first ? sd.lc[j] : sd.lw[j]) + run;
    sd.I[j] = sd.lc[j] + run;
    sd.D[j] = kImpossible;
    run += sd.lc[j];
  }
  sd.edge = run;
  int row = 1, str_right = -1;
  for (int b = 0; b < (int)hb.size(); b++) {
    const OrientedBlock& blk = hb[b];
    const int len = (int)blk.seq.size();
    if (reuse && b < last_changed) {
      row += len + (b == 0 ? -1 : 0);
      if (blk.period > 0) str_right = row - 1;
      continue;
    }
    if (blk.period > 0) {
      const int p = blk.period;
      const double* prm = loc.bt->block_stutter + 6 * (size_t)blk.src_block;
      StutterPmf pmf(prm, p);
      const int max_ins = HIPSTR_MAX_ARTIFACT_UNITS * p, max_del = -HIPSTR_MAX_ARTIFACT_UNITS * p;
      // left_align_ is !reversed_ of the owning RepeatBlock (RepeatBlock.h:28,41)
      RepeatEval ev(blk.seq, p, !reversed);
      ev.load(sd.rd.data(), sd.lc.data(), sd.lw.data(), n);
      const size_t prev = (size_t)n * (row - 1);
      size_t out = (size_t)n * (row + len - 1);
      double probs[HIPSTR_NUM_ARTIFACTS];
      for (int j = 0; j < n; j++, out++) {
        int a = 0;
        double best = kImpossible;
        sd.art_size[(size_t)n * b + j] = -10000;
        for (int D = max_del; D <= max_ins; D += p, a++) {
          int pos = -1;
          int base_len = std::min(len + D, j + 1);
          if (base_len >= 0) {
            double pr = ev.eval(base_len, j, D, pos);
            double pre = (j - base_len < 0 ? 0 : sd.M[j - base_len + prev]);
            // RepeatStutterInfo::log_prob_pcr_artifact (RepeatStutterInfo.h:53-61)
            int read_size = len + D;
            double art;
            if (D == 0) art = pmf(len, read_size);
            else if (D > 0) art = (D > max_ins ? kLargeNegative : pmf(len, read_size));
            else art = (D < max_del || read_size < 0 ? kLargeNegative : pmf(len, read_size));
            probs[a] = art + pr + pre;
          } else
            probs[a] = kImpossible;
          if (probs[a] > best) {
            sd.art_size[(size_t)n * b + j] = D;
            sd.art_pos[(size_t)n * b + j] = pos;
            best = probs[a];
          }
        }
        sd.M[out] = lse_vec(probs, HIPSTR_NUM_ARTIFACTS);
        sd.I[out] = kImpossible;
        sd.D[out] = kImpossible;
      }
      str_right = row + len - 1;
      row += len;
      continue;
    }
    for (int c = (b == 0 ? 1 : 0); c < len; c++, row++) {
      const char hc = blk.seq[c];
      int hp = std::min(kMaxHomop, std::max(homopolymer_len(hb, b, c), homopolymer_len(hb, b, std::max(0, c - 1))));
      size_t at = (size_t)n * row;
      const bool after_str = (row == str_right + 1);
      sd.M[at] = (sd.rd[0] == hc ? sd.lc[0] : sd.lw[0]);
      sd.I[at] = after_str ? kImpossible : sd.lc[0];
      sd.D[at] = after_str ? kImpossible : std::max(sd.D[at - n] + kDelToDel, sd.M[at - n] + kDelToMatch);
      at++;
      if (after_str) {
        for (int j = 1; j < n; j++, at++) {
          double e = (sd.rd[j] == hc ? sd.lc[j] : sd.lw[j]);
          sd.M[at] = e + sd.M[at - n - 1];
          sd.I[at] = kImpossible;
          sd.D[at] = kImpossible;
        }
        continue;
      }
      for (int j = 1; j < n; j++, at++) {
        double a0 = sd.I[at - 1] + t.m2i[hp];
        double a1 = sd.M[at - n - 1] + t.m2m[hp];
        double a2 = sd.D[at - n - 1] + t.m2d[hp];
        double e = (sd.rd[j] == hc ? sd.lc[j] : sd.lw[j]);
        sd.M[at] = e + std::max(a0, std::max(a1, a2));
        sd.I[at] = sd.lc[j] + std::max(sd.M[at - n - 1] + kInsToMatch, sd.I[at - 1] + kInsToIns);
        sd.D[at] = std::max(sd.M[at - n] + kDelToMatch, sd.D[at - n] + kDelToDel);
      }
    }
  }
}

void build_oriented(const Locus& loc, const int* opts, bool reversed, std::vector<OrientedBlock>& out) {
  out.clear();
  for (int k = 0; k < loc.nb; k++) {
    int b = reversed ? loc.nb - 1 - k : k;
    OrientedBlock ob;
    ob.seq = loc.option_seq(b, opts[b]);
    if (reversed) std::reverse(ob.seq.begin(), ob.seq.end());
    ob.period = loc.bt->block_period[loc.first_block + b];
    ob.opt = opts[b];
    ob.src_block = loc.first_block + b;
    run_lengths(ob);
    out.push_back(ob);
  }
}

// process_read (HapAligner.cpp:573-709) + compute_aln_logprob (:163-231) for one pooled read.
void align_pool(const Locus& loc, const char* bases, const char* quals, int len, int seed, const uint8_t* hap_mask,
                double* ll_row, int32_t* pos_row) {
  const Tables& t = T();
  std::vector<double> lw(len), lc(len);
  for (int j = 0; j < len; j++) {
    unsigned char q = static_cast<unsigned char>(quals[j]);
    lw[j] = t.qual_error[q];
    lc[j] = t.qual_correct[q];
  }
  SideDP L, R;
  L.n = seed;
  L.rd.assign(bases, bases + seed);
  L.lc.assign(lc.begin(), lc.begin() + seed);
  L.lw.assign(lw.begin(), lw.begin() + seed);
  R.n = len - seed - 1;
  R.rd.assign(bases + seed + 1, bases + len);
  std::reverse(R.rd.begin(), R.rd.end());
  R.lc.assign(lc.begin() + seed + 1, lc.end());
  R.lw.assign(lw.begin() + seed + 1, lw.end());
  std::reverse(R.lc.begin(), R.lc.end());
  std::reverse(R.lw.begin(), R.lw.end());
  for (SideDP* s : {&L, &R}) {
    size_t cells = (size_t)s->n * loc.max_rows;
    s->M.assign(cells, 0.0); s->I.assign(cells, 0.0); s->D.assign(cells, 0.0);
    s->art_size.assign((size_t)s->n * loc.nb, 0); s->art_pos.assign((size_t)s->n * loc.nb, 0);
  }
  std::vector<int> opts(loc.nb), prev_opts(loc.nb);
  std::vector<OrientedBlock> fw, rv;
  bool reuse = false;
  for (int64_t h = 0; h < loc.n_haps; h++) {
    hap_options(loc.nb, loc.nopt.data(), h, opts.data());
    int last_changed = -1;
    if (h > 0)
      for (int b = 0; b < loc.nb; b++) if (opts[b] != prev_opts[b]) last_changed = b;
    prev_opts = opts;
    if (hap_mask && !hap_mask[h]) { reuse = false; continue; }
    build_oriented(loc, opts.data(), false, fw);
    build_oriented(loc, opts.data(), true, rv);
    fill_side(L, fw, loc, reuse, last_changed, false);
    fill_side(R, rv, loc, reuse, last_changed < 0 ? -1 : loc.nb - 1 - last_changed, true);
    reuse = true;

    int hs = 0, num_seeds = 0;
    for (auto& b : fw) { hs += (int)b.seq.size(); if (b.period == 0) num_seeds += (int)b.seq.size(); }
    const int lf = L.n, rf = R.n;
    const double prior = -t.int_logs[num_seeds];
    const char sc = bases[seed];
    const double s_ok = lc[seed], s_bad = lw[seed];
    std::vector<double> terms;
    terms.push_back(prior + (sc == fw.front().seq[0] ? s_ok : s_bad) + L.edge + R.M[(size_t)rf * (hs - 1) - 1]);
    int best_pos = 0;
    double best = terms[0];
    terms.push_back(prior + (sc == fw.back().seq.back() ? s_ok : s_bad) + R.edge + L.M[(size_t)lf * (hs - 1) - 1]);
    if (terms[1] > best) { best_pos = hs - 1; best = terms[1]; }
    int hp = 1;
    for (int b = 0; b < loc.nb; b++) {
      const std::string& s = fw[b].seq;
      if (fw[b].period > 0) { hp += (int)s.size(); continue; }
      int c0 = (b == 0 ? 1 : 0), c1 = (b == loc.nb - 1 ? (int)s.size() - 1 : (int)s.size());
      for (int c = c0; c < c1; c++, hp++) {
        double v = prior + (sc == s[c] ? s_ok : s_bad) + L.M[(size_t)lf * hp - 1] + R.M[(size_t)rf * (hs - hp - 1) - 1];
        terms.push_back(v);
        if (v > best) { best_pos = hp; best = v; }
      }
    }
    ll_row[h] = lse_vec(terms.data(), (int)terms.size());
    if (pos_row) pos_row[h] = best_pos;
  }
}

}  // namespace


// ---------------------------------------------------------------------------
// EM stutter learner (em_stutter_genotyper.cpp:10-226, em_stutter_genotyper.h:50-101),
// one locus.  Same loops, same accumulation orders, same libm.
// ---------------------------------------------------------------------------
namespace {

double exact_lse2(double a, double b) {                       // mathops.cpp:51-56
  return a > b ? a + std::log(1 + std::exp(b - a)) : b + std::log(1 + std::exp(a - b));
}
double exact_lse3(double a, double b, double c) {             // mathops.cpp:58-61
  double mx = std::max(std::max(a, b), c);
  return mx + std::log(std::exp(a - mx) + std::exp(b - mx) + std::exp(c - mx));
}
void stream_lse(double v, double& mx, double& total) {        // mathops.cpp:72-80
  if (v <= mx) total += std::exp(v - mx);
  else { total *= std::exp(mx - v); total += 1.0; mx = v; }
}

struct EmLocus {
  int R, S, A, period;
  bool haploid;
  std::vector<int> bps, allele_of, label, per_sample;
  const double *p1, *p2;
  std::vector<double> gt_prior, post, sll;
  double prm[6];

  double prior(int a, int b) const {                           // em_stutter_genotyper.cpp:129-144 (use_pop_freqs_)
    if (!haploid) return gt_prior[a] + gt_prior[b];
    return a == b ? gt_prior[a] : -DBL_MAX / 2;
  }
  // E-step: calc_hap_aln_probs + calc_log_sample_posteriors (genotyper.cpp:44-80)
  double posteriors(const StutterPmf& pmf) {
    const Tables& t = T();
    for (int s = 0; s < S; s++)
      for (int a = 0; a < A; a++)
        for (int b = 0; b < A; b++) post[((size_t)s * A + a) * A + b] = prior(a, b);
    std::vector<double> row(A);
    for (int r = 0; r < R; r++) {
      for (int a = 0; a < A; a++) row[a] = pmf(bps[a], bps[allele_of[r]]);
      double* sp = &post[(size_t)label[r] * A * A];
      for (int a = 0; a < A; a++)
        for (int b = 0; b < A; b++, sp++) *sp += 1 * lse2(t.log_half + p1[r] + row[a], t.log_half + p2[r] + row[b]);
    }
    double total = 0.0;
    for (int s = 0; s < S; s++) {
      double* sp = &post[(size_t)s * A * A];
      sll[s] = exact_lse(sp, sp + (size_t)A * A);
      for (int i = 0; i < A * A; i++) sp[i] -= sll[s];
    }
    for (int s = 0; s < S; s++) total += sll[s];
    return total;
  }
  void recalc_gt_priors() {                                    // :21-56
    std::vector<double> mx(A, -DBL_MAX / 2), tot(A, 0.0);
    const double* p = post.data();
    for (int s = 0; s < S; s++)
      for (int a = 0; a < A; a++, p += A) stream_lse(exact_lse(p, p + A), mx[a], tot[a]);
    p = post.data();
    for (int s = 0; s < S; s++)
      for (int a = 0; a < A; a++)
        for (int b = 0; b < A; b++, p++) stream_lse(*p, mx[b], tot[b]);
    for (int a = 0; a < A; a++) gt_prior[a] = mx[a] + std::log(tot[a]);
    double lt = exact_lse(gt_prior.data(), gt_prior.data() + A);
    for (int a = 0; a < A; a++) gt_prior[a] -= lt;
  }
  void recalc_model(const StutterPmf& pmf) {                   // :63-127 with :152-168 folded in
    const Tables& t = T();
    std::vector<double> in_up(1, 0.0), in_down(1, 0.0), in_eq(1, 0.0), in_diffs, out_up(1, 0.0), out_down(1, 0.0), out_diffs;
    in_diffs.push_back(0.0); in_diffs.push_back(std::log(1.1));
    out_diffs.push_back(0.0); out_diffs.push_back(std::log(1.1));
    for (int r = 0; r < R; r++) {
      const double* gp = &post[(size_t)label[r] * A * A];
      const int rb = bps[allele_of[r]];
      for (int a = 0; a < A; a++)
        for (int b = 0; b < A; b++, gp++) {
          double one = t.log_half + p1[r] + pmf(bps[a], rb), two = t.log_half + p2[r] + pmf(bps[b], rb);
          double both = lse2(one, two);
          double phase[2] = {one - both, two - both};
          for (int ph = 0; ph < 2; ph++) {
            int gt = ph == 0 ? a : b, d = rb - bps[gt];
            double f = *gp + phase[ph];
            if (d == 0) in_eq.push_back(f);
            else if (d % period != 0) {
              int eff = d - d / period;
              out_diffs.push_back(f + t.int_logs[std::abs(eff)]);
              (d > 0 ? out_up : out_down).push_back(f);
            } else {
              int eff = d / period;
              in_diffs.push_back(f + t.int_logs[std::abs(eff)]);
              (d > 0 ? in_up : in_down).push_back(f);
            }
          }
        }
    }
    auto L = [](std::vector<double>& v) { return lse_vec(v.data(), (int)v.size()); };
    double iu = L(in_up), id = L(in_down), ie = L(in_eq), idf = L(in_diffs), ou = L(out_up), od = L(out_down), odf = L(out_diffs);
    double ot = lse2(ou, od);
    prm[0] = std::min(0.999, std::exp(exact_lse2(iu, id) - idf));
    prm[3] = std::min(0.999, std::exp(ot - odf));
    double lt = exact_lse2(exact_lse3(iu, id, ie), ot);
    prm[1] = std::exp(iu - lt); prm[2] = std::exp(id - lt);
    prm[4] = std::exp(ou - lt); prm[5] = std::exp(od - lt);
  }
};

}  // namespace

extern "C" int32_t oracle_em_train(const hipstr_em_batch_t* bt, int32_t max_iter, double min_abs, double min_frac,
                                   double* params_out, uint8_t* converged_out, int32_t* iters_out, double* ll_out) {
  for (int l = 0; l < bt->n_loci; l++) {
    EmLocus e;
    const int r0 = bt->locus_read_off[l], r1 = bt->locus_read_off[l + 1];
    e.R = r1 - r0; e.S = bt->locus_sample_off[l + 1] - bt->locus_sample_off[l];
    e.period = bt->motif_len[l]; e.haploid = bt->haploid[l] != 0;
    e.p1 = bt->log_p1 + r0; e.p2 = bt->log_p2 + r0;
    // allele list: reference first, the rest sorted (em_stutter_genotyper.h:59-81)
    std::vector<int> sizes(bt->num_bps + r0, bt->num_bps + r1);
    std::sort(sizes.begin(), sizes.end());
    sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
    sizes.erase(std::remove(sizes.begin(), sizes.end(), bt->ref_allele[l]), sizes.end());
    e.bps.push_back(bt->ref_allele[l]);
    e.bps.insert(e.bps.end(), sizes.begin(), sizes.end());
    e.A = (int)e.bps.size();
    e.per_sample.assign(e.S, 0);
    for (int r = r0; r < r1; r++) {
      e.label.push_back(bt->sample_label[r]);
      e.per_sample[bt->sample_label[r]]++;
      e.allele_of.push_back((int)(std::find(e.bps.begin(), e.bps.end(), bt->num_bps[r]) - e.bps.begin()));
    }
    e.post.assign((size_t)e.S * e.A * e.A, 0.0);
    e.sll.assign(e.S, 0.0);
    // init_log_gt_priors (:10-19)
    e.gt_prior.assign(e.A, 1.0);
    for (int r = 0; r < e.R; r++) e.gt_prior[e.allele_of[r]] += 1.0 / e.per_sample[e.label[r]];
    double tot = 0.0;
    for (int a = 0; a < e.A; a++) tot += e.gt_prior[a];
    const double log_total = std::log(tot);
    for (int a = 0; a < e.A; a++) e.gt_prior[a] = std::log(e.gt_prior[a]) - log_total;
    const double init[6] = {0.9, 0.1, 0.1, 0.8, 0.01, 0.01};   // :58-61
    std::copy(init, init + 6, e.prm);
    // train (:170-226)
    int iter = 1;
    double LL = -DBL_MAX;
    bool converged = false;
    while (iter <= max_iter) {
      StutterPmf pmf(e.prm, e.period);
      double new_LL = e.posteriors(pmf);
      if (new_LL < LL + 1e-10) { converged = true; LL = new_LL; break; }
      e.recalc_gt_priors();
      double prev[6];
      std::copy(e.prm, e.prm + 6, prev);
      e.recalc_model(pmf);
      const double abs_change = new_LL - LL, frac_change = -(new_LL - LL) / LL;
      bool close = true;
      for (int k = 0; k < 6; k++) close = close && std::fabs(prev[k] - e.prm[k]) < 0.0001;
      LL = new_LL;
      if ((abs_change < min_abs && frac_change < min_frac) || close) { converged = true; break; }
      iter++;
    }
    std::copy(e.prm, e.prm + 6, params_out + 6 * (size_t)l);
    converged_out[l] = converged;
    if (iters_out) iters_out[l] = std::min(iter, max_iter);
    if (ll_out) ll_out[l] = LL;
  }
  return HIPSTR_OK;
}

// genotyper.cpp:129-251 (+ calc_PLs :99-104, calc_gl_diff :106-127), one locus at a time
extern "C" int32_t oracle_extract_genotypes(int32_t n_loci, const int32_t* locus_sample_off, const int32_t* n_haps,
                                            const int32_t* n_variants, const int32_t* hap_to_allele, const uint8_t* haploid,
                                            const double* post, const double* sample_ll, int32_t* best_hap, int32_t* best_gt,
                                            double* log_phased, double* log_unphased, double* hap_log_phased,
                                            double* hap_log_unphased, double* gl, double* phased_gl, double* gl_diff,
                                            int32_t* pl) {
  const Tables& t = T();
  const double LOG_E_BASE_10 = 0.4342944819;   // mathops.cpp:11
  size_t post_off = 0, h2a_off = 0, gl_off = 0, pgl_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l], V = n_variants[l], s0 = locus_sample_off[l], S = locus_sample_off[l + 1] - s0;
    const bool hap1 = haploid[l] != 0;
    const int32_t* h2a = hap_to_allele + h2a_off;
    const int G = hap1 ? V : V * (V + 1) / 2, PG = hap1 ? V : V * V;
    const double hom = hap1 ? -t.int_logs[H] : t.int_logs[2] - t.int_logs[H] - t.int_logs[H + 1];
    const double het = hap1 ? 0 : -t.int_logs[H] - t.int_logs[H + 1];
    const double gl_nconfig = hap1 ? t.int_logs[2] + t.int_logs[H] - t.int_logs[V] : t.int_logs[2] + 2 * (t.int_logs[H] - t.int_logs[V]);
    const double pgl_nconfig = hap1 ? t.int_logs[H] - t.int_logs[V] : 2 * (t.int_logs[H] - t.int_logs[V]);
    for (int s = 0; s < S; s++) {
      const double* sp = post + post_off + (size_t)s * H * H;
      // get_optimal_haplotypes (:82-97)
      double best = -DBL_MAX; int ba = -1, bb = -1;
      for (int a = 0; a < H; a++) for (int b = 0; b < H; b++) if (sp[a * H + b] > best) { best = sp[a * H + b]; ba = a; bb = b; }
      best_hap[2 * (s0 + s)] = ba; best_hap[2 * (s0 + s) + 1] = bb;
      const int ga = h2a[ba], gb = h2a[bb];
      best_gt[2 * (s0 + s)] = ga; best_gt[2 * (s0 + s) + 1] = gb;
      // streaming marginalisation (:152-171)
      std::vector<double> mx((size_t)V * V, -DBL_MAX / 2), tot((size_t)V * V, 0.0);
      for (int a = 0; a < H; a++) for (int b = 0; b < H; b++) stream_lse(sp[a * H + b], mx[V * h2a[a] + h2a[b]], tot[V * h2a[a] + h2a[b]]);
      for (int g = 0; g < V * V; g++) tot[g] = mx[g] + std::log(tot[g]);
      hap_log_phased[s0 + s] = sp[ba * H + bb];
      hap_log_unphased[s0 + s] = ba != bb ? lse2(sp[ba * H + bb], sp[bb * H + ba]) : sp[ba * H + bb];
      const double lp = tot[V * ga + gb];
      log_phased[s0 + s] = lp;
      log_unphased[s0 + s] = ga == gb ? lp : exact_lse2(lp, tot[V * gb + ga]);
      double* g_out = gl + gl_off + (size_t)s * G;
      double* pg_out = phased_gl + pgl_off + (size_t)s * PG;
      int gi = 0, pi = 0;
      for (int i1 = 0; i1 < V; i1++)
        for (int i2 = 0; i2 < V; i2++) {
          const int g = i1 * V + i2, ag = i2 * V + i1;
          const double glc = (i1 == i2 ? hom : het) + gl_nconfig, pglc = (i1 == i2 ? hom : het) + pgl_nconfig;
          if (i2 <= i1 && (!hap1 || i1 == i2)) g_out[gi++] = (sample_ll[s0 + s] - glc + lse2(tot[g], tot[ag])) * LOG_E_BASE_10;
          if (!hap1 || i1 == i2) pg_out[pi++] = (sample_ll[s0 + s] - pglc + tot[g]) * LOG_E_BASE_10;
        }
      // calc_gl_diff (:106-127)
      double d;
      if (H == 1) d = -1000;
      else {
        double mg = g_out[0];
        for (int i = 1; i < G; i++) mg = std::max(mg, g_out[i]);
        double second = -DBL_MAX;
        for (int i = 0; i < G; i++) if (g_out[i] < mg) second = std::max(second, g_out[i]);
        if (second == -DBL_MAX) second = mg;
        const int idx = hap1 ? ga : std::max(ga, gb) * (std::max(ga, gb) + 1) / 2 + std::min(ga, gb);
        d = std::fabs(mg - g_out[idx]) < 1e-10 ? mg - second : g_out[idx] - mg;
      }
      gl_diff[s0 + s] = d;
      double mg = g_out[0];
      for (int i = 1; i < G; i++) mg = std::max(mg, g_out[i]);
      int32_t* pl_out = pl + gl_off + (size_t)s * G;
      for (int i = 0; i < G; i++) pl_out[i] = std::min(999, (int)(-10 * (g_out[i] - mg)));
    }
    post_off += (size_t)S * H * H; h2a_off += H; gl_off += (size_t)S * G; pgl_off += (size_t)S * PG;
  }
  return HIPSTR_OK;
}

// ---------------------------------------------------------------------------
// Traceback: HapAligner::process_read with retrace_aln = true for ONE haplotype
// (HapAligner.cpp:636-690) and HapAligner::retrace (:363-571).
// ---------------------------------------------------------------------------
namespace {

const double kTraceTol = 0.001;                    // HapAligner.cpp:345
const double kMinSnpLogCorrect = -0.0043648054;    // HapAligner.cpp:24

// tie-break helpers (:346-361); `rev` selects the mirrored preference order
int pick3(bool rev, double v1, double v2, double v3) {
  if (!rev) {
    if (v1 > v2 + kTraceTol) return v1 > v3 + kTraceTol ? 0 : 2;
    return v2 > v3 + kTraceTol ? 1 : 2;
  }
  if (v3 > v2 + kTraceTol) return v3 > v1 + kTraceTol ? 2 : 0;
  return v2 > v1 + kTraceTol ? 1 : 0;
}
int pick2(bool rev, double v1, double v2) {
  if (!rev) return v1 > v2 + kTraceTol ? 0 : 1;
  return v2 > v1 + kTraceTol ? 1 : 0;
}

struct TraceAcc {
  int nb;
  std::vector<int> stutter, lo, hi;   // per FORWARD block; lo > hi = nothing recorded
  int ins = 0, del = 0;
  std::vector<std::pair<int, int> > indels, snps;
  explicit TraceAcc(int n) : nb(n), stutter(n, HIPSTR_NO_STR_DATA), lo(n, 1 << 30), hi(n, -1) {}
  void touch(int block, int read_index) { lo[block] = std::min(lo[block], read_index); hi[block] = std::max(hi[block], read_index); }
};

// One side.  `side_to_read(k)` maps a column of this side to the index of that base in the read.
std::string walk_back(const std::vector<OrientedBlock>& hb, bool rev, const SideDP& sd, const std::vector<int>& start_of,
                      int block_index, int base_index, long matrix_index, int read_len, TraceAcc& acc) {
  const Tables& t = T();
  const int n = sd.n, nb = (int)hb.size();
  auto to_read = [&](int k) { return rev ? read_len - 1 - k : k; };
  int seq_index = n - 1, type = 0;   // 0 MATCH, 1 DEL, 2 INS
  std::string aln;
  while (block_index >= 0) {
    const OrientedBlock& blk = hb[block_index];
    const int fw_block = rev ? nb - 1 - block_index : block_index;
    if (blk.period > 0) {
      const int size = sd.art_size[(size_t)n * block_index + seq_index], pos = sd.art_pos[(size_t)n * block_index + seq_index];
      const int len = (int)blk.seq.size();
      int i = 0;
      for (; i < std::min(seq_index + 1, pos); i++) { aln += 'M'; acc.touch(fw_block, to_read(seq_index - i)); }
      if (size < 0) aln += std::string(-size, 'D');
      else for (; i < std::min(seq_index + 1, pos + size); i++) { aln += 'I'; acc.touch(fw_block, to_read(seq_index - i)); }
      for (; i < std::min(len + size, seq_index + 1); i++) { aln += 'M'; acc.touch(fw_block, to_read(seq_index - i)); }
      acc.stutter[fw_block] = size;
      if (len + size >= seq_index + 1) return aln;   // the read ends inside the repeat block
      matrix_index -= (len + size + (long)n * len);
      type = 0;
      seq_index -= (len + size);
    } else {
      int prev_type = -1;
      int pos = start_of[block_index] + (rev ? -base_index : base_index);
      const int step = rev ? 1 : -1;
      int indel_seq_index = -1, indel_pos = -1;
      while (base_index >= 0 && seq_index >= 0) {
        const int hp = std::min(kMaxHomop, std::max(homopolymer_len(hb, block_index, base_index),
                                                    homopolymer_len(hb, block_index, std::max(0, base_index - 1))));
        if (type != prev_type) {
          if (prev_type == 1) acc.indels.push_back(rev ? std::make_pair(indel_pos, indel_pos - pos) : std::make_pair(pos + 1, pos - indel_pos));
          else if (prev_type == 2) acc.indels.push_back(std::make_pair(indel_pos + (rev ? 0 : 1), indel_seq_index - seq_index));
          if (type == 1 || type == 2) { indel_seq_index = seq_index; indel_pos = pos; }
          prev_type = type;
        }
        if (type == 0) {
          if (blk.seq[base_index] != sd.rd[seq_index] && sd.lc[seq_index] > kMinSnpLogCorrect) acc.snps.push_back(std::make_pair(pos, (int)sd.rd[seq_index]));
          acc.touch(fw_block, to_read(seq_index));
          aln += 'M'; seq_index--; base_index--; pos += step;
        } else if (type == 1) {
          acc.del++; aln += 'D'; base_index--; pos += step;
        } else {
          acc.ins++; acc.touch(fw_block, to_read(seq_index)); aln += 'I'; seq_index--;
        }
        if (seq_index == -1 || (base_index == -1 && block_index == 0)) {
          for (; seq_index != -1; seq_index--) aln += 'S';
          return aln;
        }
        if (type == 0) {
          const int best = pick3(rev, sd.I[matrix_index - 1] + t.m2i[hp], sd.D[matrix_index - n - 1] + t.m2d[hp], sd.M[matrix_index - n - 1] + t.m2m[hp]);
          if (best == 0) { type = 2; matrix_index -= 1; }
          else { type = best == 1 ? 1 : 0; matrix_index -= n + 1; }
        } else if (type == 1) {
          type = pick2(rev, sd.D[matrix_index - n] + kDelToDel, sd.M[matrix_index - n] + kDelToMatch) == 0 ? 1 : 0;
          matrix_index -= n;
        } else {
          if (pick2(rev, sd.I[matrix_index - 1] + kInsToIns, sd.M[matrix_index - n - 1] + kInsToMatch) == 0) { type = 2; matrix_index -= 1; }
          else { type = 0; matrix_index -= n + 1; }
        }
      }
    }
    --block_index;
    if (block_index >= 0) base_index = (int)hb[block_index].seq.size() - 1;
  }
  return aln;
}

}  // namespace

extern "C" int32_t oracle_trace_batch(const hipstr_align_batch_t* bt, const int32_t* block_start, int32_t n_traces,
                                      const int32_t* trace_pool, const int32_t* trace_hap, const hipstr_trace_out_t* out) {
  const Tables& t = T();
  for (int tr = 0; tr < n_traces; tr++) {
    const int p = trace_pool[tr];
    int l = 0;
    while (!(p >= bt->locus_pool_off[l] && p < bt->locus_pool_off[l + 1])) l++;
    Locus loc(bt, l);
    const int nb = loc.nb;
    const int s0 = bt->pool_seq_off[p], len = bt->pool_seq_off[p + 1] - s0, seed = bt->pool_seed[p];
    if (seed <= 0 || seed >= len - 1 || trace_hap[tr] < 0 || trace_hap[tr] >= loc.n_haps) return HIPSTR_ERR_BAD_ARG;
    const char* bases = bt->pool_bases + s0;
    const char* quals = bt->pool_quals + s0;
    std::vector<double> lw(len), lc(len);
    for (int j = 0; j < len; j++) { unsigned char q = (unsigned char)quals[j]; lw[j] = t.qual_error[q]; lc[j] = t.qual_correct[q]; }
    SideDP L, R;
    L.n = seed; L.rd.assign(bases, bases + seed); L.lc.assign(lc.begin(), lc.begin() + seed); L.lw.assign(lw.begin(), lw.begin() + seed);
    R.n = len - seed - 1; R.rd.assign(bases + seed + 1, bases + len); std::reverse(R.rd.begin(), R.rd.end());
    R.lc.assign(lc.begin() + seed + 1, lc.end()); R.lw.assign(lw.begin() + seed + 1, lw.end());
    std::reverse(R.lc.begin(), R.lc.end()); std::reverse(R.lw.begin(), R.lw.end());
    for (SideDP* sd : {&L, &R}) {
      size_t cells = (size_t)sd->n * loc.max_rows;
      sd->M.assign(cells, 0.0); sd->I.assign(cells, 0.0); sd->D.assign(cells, 0.0);
      sd->art_size.assign((size_t)sd->n * nb, 0); sd->art_pos.assign((size_t)sd->n * nb, 0);
    }
    std::vector<int> opts(nb);
    hap_options(nb, loc.nopt.data(), trace_hap[tr], opts.data());
    std::vector<OrientedBlock> fw, rv;
    build_oriented(loc, opts.data(), false, fw);
    build_oriented(loc, opts.data(), true, rv);
    fill_side(L, fw, loc, false, -1, false);
    fill_side(R, rv, loc, false, -1, true);
    // best seed placement (compute_aln_logprob, :163-231)
    int hs = 0, num_seeds = 0;
    for (auto& b : fw) { hs += (int)b.seq.size(); if (b.period == 0) num_seeds += (int)b.seq.size(); }
    const int lf = L.n, rf = R.n;
    const double prior = -t.int_logs[num_seeds];
    const char sc = bases[seed];
    int max_index = 0;
    double best = prior + (sc == fw.front().seq[0] ? lc[seed] : lw[seed]) + L.edge + R.M[(size_t)rf * (hs - 1) - 1];
    {
      const double v = prior + (sc == fw.back().seq.back() ? lc[seed] : lw[seed]) + R.edge + L.M[(size_t)lf * (hs - 1) - 1];
      if (v > best) { max_index = hs - 1; best = v; }
      int hp = 1;
      for (int b = 0; b < nb; b++) {
        const std::string& sq = fw[b].seq;
        if (fw[b].period > 0) { hp += (int)sq.size(); continue; }
        int c0 = (b == 0 ? 1 : 0), c1 = (b == nb - 1 ? (int)sq.size() - 1 : (int)sq.size());
        for (int c = c0; c < c1; c++, hp++) {
          const double w = prior + (sc == sq[c] ? lc[seed] : lw[seed]) + L.M[(size_t)lf * hp - 1] + R.M[(size_t)rf * (hs - hp - 1) - 1];
          if (w > best) { max_index = hp; best = w; }
        }
      }
    }
    // genomic start() of the oriented blocks: forward = block start; reversed block = end - 1 of the original
    std::vector<int> start_fw(nb), start_rv(nb);
    for (int b = 0; b < nb; b++) {
      start_fw[b] = block_start[loc.first_block + b];
      start_rv[nb - 1 - b] = block_start[loc.first_block + b] + loc.option_len(b, 0) - 1;
    }
    auto coords = [&](const std::vector<OrientedBlock>& hb, int pos, int& blk, int& off) {
      for (blk = 0; pos >= (int)hb[blk].seq.size(); blk++) pos -= (int)hb[blk].seq.size();
      off = pos;
    };
    TraceAcc acc(nb);
    std::string left, right;
    int fb, fc;
    coords(fw, max_index, fb, fc);
    if (max_index == 0) left = std::string(seed, 'S');
    else {
      const long mi = (long)seed * max_index - 1;
      if (fc == 0) left = walk_back(fw, false, L, start_fw, fb - 1, (int)fw[fb - 1].seq.size() - 1, mi, len, acc);
      else left = walk_back(fw, false, L, start_fw, fb, fc - 1, mi, len, acc);
    }
    std::reverse(left.begin(), left.end());
    if (fw[fb].period == 0) acc.touch(fb, seed);
    const int rmax = hs - 1 - max_index;
    int rb, rc;
    coords(rv, rmax, rb, rc);
    if (rmax == 0) right = std::string(len - 1 - seed, 'S');
    else {
      const long mi = (long)(len - 1 - seed) * rmax - 1;
      if (rc == 0) right = walk_back(rv, true, R, start_rv, rb - 1, (int)rv[rb - 1].seq.size() - 1, mi, len, acc);
      else right = walk_back(rv, true, R, start_rv, rb, rc - 1, mi, len, acc);
    }
    const std::string aln = left + "M" + right;
    if ((int)aln.size() + 1 > out->aln_stride) return HIPSTR_ERR_BAD_ARG;
    std::memcpy(out->hap_aln + (size_t)tr * out->aln_stride, aln.c_str(), aln.size() + 1);
    out->seed_hap_pos[tr] = max_index;
    for (int b = 0; b < HIPSTR_MAX_BLOCKS_PER_LOCUS; b++) {
      const size_t o = (size_t)tr * HIPSTR_MAX_BLOCKS_PER_LOCUS + b;
      out->stutter_size[o] = b < nb ? acc.stutter[b] : HIPSTR_NO_STR_DATA;
      const bool any = b < nb && acc.hi[b] >= acc.lo[b];
      out->span_start[o] = any ? acc.lo[b] : 0;
      out->span_len[o] = any ? acc.hi[b] - acc.lo[b] + 1 : 0;
    }
    out->flank_ins[tr] = acc.ins; out->flank_del[tr] = acc.del;
    out->n_indels[tr] = (int)acc.indels.size(); out->n_snps[tr] = (int)acc.snps.size();
    for (int k = 0; k < HIPSTR_MAX_TRACE_INDELS; k++) {
      const size_t o = ((size_t)tr * HIPSTR_MAX_TRACE_INDELS + k) * 2;
      out->indels[o] = k < (int)acc.indels.size() ? acc.indels[k].first : 0;
      out->indels[o + 1] = k < (int)acc.indels.size() ? acc.indels[k].second : 0;
    }
    for (int k = 0; k < HIPSTR_MAX_TRACE_SNPS; k++) {
      const size_t o = ((size_t)tr * HIPSTR_MAX_TRACE_SNPS + k) * 2;
      out->snps[o] = k < (int)acc.snps.size() ? acc.snps[k].first : 0;
      out->snps[o + 1] = k < (int)acc.snps.size() ? acc.snps[k].second : 0;
    }
  }
  return HIPSTR_OK;
}

extern "C" {

double oracle_fast_lse2(double a, double b) { return lse2(a, b); }
double oracle_fast_lse_vec(const double* v, int32_t n) { return lse_vec(v, n); }

void oracle_hap_options(int32_t n_blocks, const int32_t* n_opts, int64_t hap, int32_t* out_opts) {
  hap_options(n_blocks, n_opts, hap, out_opts);
}

// calc_seed_base (HapAligner.cpp:270-318) + calc_best_seed_position (:238-264)
int32_t oracle_calc_seeds(int32_t n_reads, const int32_t* read_start, const int32_t* read_len,
                          const int32_t* cigar_off, const char* cigar_type, const int32_t* cigar_len,
                          int32_t first_block_start, int32_t last_block_end, int32_t n_repeats,
                          const int32_t* repeat_start, const int32_t* repeat_end, int32_t* out_seed) {
  for (int r = 0; r < n_reads; r++) {
    int32_t pos = read_start[r];
    int best_seed = -1, cur_base = 0, max_dist = kMinSeedDist;
    for (int c = cigar_off[r]; c < cigar_off[r + 1]; c++) {
      const int num = cigar_len[c];
      switch (cigar_type[c]) {
        case '=': {
          int32_t lo = std::max(pos, first_block_start), hi = std::min(pos + num - 1, last_block_end - 1);
          if (lo <= hi) {
            int32_t bd = -1, bp = -1, at = lo;
            int ri = 0;
            while (ri < n_repeats && at <= hi) {
              if (at < repeat_start[ri]) {
                int32_t d = 1 + (std::min(hi, repeat_start[ri] - 1) - at) / 2;
                if (d >= bd) { bd = d; bp = d - 1 + at; }
                at = repeat_end[ri++];
              } else if (at < repeat_end[ri])
                at = repeat_end[ri++];
              else
                ri++;
            }
            if (at <= hi) {
              int32_t d = 1 + (hi - at) / 2;
              if (d >= bd) { bd = d; bp = d - 1 + at; }
            }
            if (bd >= max_dist) { max_dist = bd; best_seed = cur_base + (bp - pos); }
          }
          pos += num; cur_base += num;
          break;
        }
        case 'I': cur_base += num; break;
        case 'X': pos += num; cur_base += num; break;
        case 'D': pos += num; break;
        default: return HIPSTR_ERR_BAD_CIGAR;
      }
    }
    if (best_seed < -1 || best_seed == 0 || best_seed >= read_len[r] - 1) return HIPSTR_ERR_INVALID_SEED;
    out_seed[r] = best_seed;
  }
  return HIPSTR_OK;
}

int32_t oracle_align_loci(const hipstr_align_batch_t* bt, int32_t l0, int32_t l1, double* ll_out,
                          int32_t* seed_hap_pos) {
  for (int l = l0; l < l1; l++) {
    Locus loc(bt, l);
    const uint8_t* hmask = bt->realign_hap ? bt->realign_hap + bt->locus_hap_off[l] : nullptr;
    int p0 = bt->locus_pool_off[l], p1 = bt->locus_pool_off[l + 1];
    for (int p = p0; p < p1; p++) {
      if (bt->realign_pool && !bt->realign_pool[p]) continue;
      double* row = ll_out + bt->locus_out_off[l] + (int64_t)(p - p0) * loc.n_haps;
      int32_t* prow = seed_hap_pos ? seed_hap_pos + bt->locus_out_off[l] + (int64_t)(p - p0) * loc.n_haps : nullptr;
      int seed = bt->pool_seed[p];
      if (seed < 0) {   // HapAligner.cpp:333-337: every haplotype, mask ignored
        for (int64_t h = 0; h < loc.n_haps; h++) row[h] = 0;
        continue;
      }
      int s0 = bt->pool_seq_off[p], len = bt->pool_seq_off[p + 1] - s0;
      align_pool(loc, bt->pool_bases + s0, bt->pool_quals + s0, len, seed, hmask, row, prow);
    }
  }
  return HIPSTR_OK;
}

int32_t oracle_align_batch(const hipstr_align_batch_t* bt, double* ll_out, int32_t* seed_hap_pos) {
  return oracle_align_loci(bt, 0, bt->n_loci, ll_out, seed_hap_pos);
}

// seq_stutter_genotyper.cpp:530-564
int32_t oracle_scatter_pool_lls(int32_t n_reads, int32_t n_haps, const double* pool_ll, const int32_t* pool_seed,
                                const int32_t* pool_index, const uint8_t* second_mate, const uint8_t* copy_read,
                                const uint8_t* realign_hap, double* read_ll, int32_t* read_seed) {
  for (int r = 0; r < n_reads; r++) {
    if (copy_read && !copy_read[r]) continue;
    if (read_seed) read_seed[r] = pool_seed[pool_index[r]];
    for (int h = 0; h < n_haps; h++)
      if (!realign_hap || realign_hap[h]) read_ll[(size_t)r * n_haps + h] = pool_ll[(size_t)pool_index[r] * n_haps + h];
  }
  for (int r = 0; r < n_reads; r++) {
    if (!second_mate[r] || (copy_read && !copy_read[r])) continue;
    for (int h = 0; h < n_haps; h++)
      if (!realign_hap || realign_hap[h]) {
        double total = read_ll[(size_t)(r - 1) * n_haps + h] + read_ll[(size_t)r * n_haps + h];
        read_ll[(size_t)(r - 1) * n_haps + h] = total;
        read_ll[(size_t)r * n_haps + h] = total;
      }
  }
  return HIPSTR_OK;
}

// genotyper.cpp:20-97
int32_t oracle_posteriors(int32_t n_loci, const int32_t* locus_read_off, const int32_t* locus_sample_off,
                          const int32_t* n_haps, const uint8_t* haploid, const double* read_ll,
                          const double* log_p1, const double* log_p2, const int32_t* sample_label,
                          const int32_t* read_weight, double* post_out, double* sample_ll_out, int32_t* best_out,
                          double* total_ll_out) {
  const Tables& t = T();
  size_t ll_off = 0, post_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l], r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    const int s0 = locus_sample_off[l], S = locus_sample_off[l + 1] - s0;
    double homoz, hetz;
    if (haploid[l]) { homoz = -t.int_logs[H]; hetz = -DBL_MAX / 2; }
    else { homoz = t.int_logs[2] - t.int_logs[H] - t.int_logs[H + 1]; hetz = -t.int_logs[H] - t.int_logs[H + 1]; }
    double* post = post_out + post_off;
    for (int s = 0; s < S; s++)
      for (int a = 0; a < H; a++)
        for (int b = 0; b < H; b++) post[((size_t)s * H + a) * H + b] = (a == b ? homoz : hetz);
    const double* ll = read_ll + ll_off;
    for (int r = r0; r < r1; r++, ll += H) {
      double* sp = post + (size_t)sample_label[r] * H * H;
      for (int a = 0; a < H; a++)
        for (int b = 0; b < H; b++, sp++)
          *sp += read_weight[r] * lse2(t.log_half + log_p1[r] + ll[a], t.log_half + log_p2[r] + ll[b]);
    }
    double total = 0.0;
    for (int s = 0; s < S; s++) {
      double* sp = post + (size_t)s * H * H;
      double sll = exact_lse(sp, sp + (size_t)H * H);
      sample_ll_out[s0 + s] = sll;
      for (int i = 0; i < H * H; i++) sp[i] -= sll;
      total += sll;
      if (best_out) {
        double best = -DBL_MAX;
        int ba = -1, bb = -1;
        for (int a = 0; a < H; a++)
          for (int b = 0; b < H; b++)
            if (sp[a * H + b] > best) { best = sp[a * H + b]; ba = a; bb = b; }
        best_out[2 * (s0 + s)] = ba;
        best_out[2 * (s0 + s) + 1] = bb;
      }
    }
    if (total_ll_out) total_ll_out[l] = total;
    ll_off += (size_t)(r1 - r0) * H;
    post_off += (size_t)S * H * H;
  }
  return HIPSTR_OK;
}

/* NeedlemanWunsch::Align (SeqAlignment/NeedlemanWunsch.cpp:384-423) for one (reference window, read) pair, serial,
 * with the three full matrices: the checker of kernel K6.  ops: 'M' base vs base, 'D' reference base vs gap,
 * 'I' read base vs gap.  Returns the number of columns or -1. */
int32_t oracle_nw_align(const char* ref, int32_t L1, const char* read, int32_t L2, int32_t use_ref_end_penalty, char* ops,
                        float* score) {
  const float kOpen = 5.0f, kExt = 0.125f, kLarge = 1000000.0f;
  auto code = [](char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; default: return 4; }
  };
  auto pick = [](float s1, float s2, float s3, int8_t* w) {   // bestIndex :125-147
    if (s2 > s1) { if (s2 > s3) { *w = 1; return s2; } *w = 2; return s3; }
    if (s3 > s1) { *w = 2; return s3; }
    *w = 0; return s1;
  };
  const int W = L1 + 1;
  const size_t cells = (size_t)W * (L2 + 1);
  std::vector<float> M(cells), X(cells), Y(cells);
  std::vector<int8_t> tM(cells, -1), tX(cells, -1), tY(cells, -1);
  M[0] = 0.0f; X[0] = -kLarge; Y[0] = -kLarge;
  for (int j = 1; j <= L1; j++) { X[j] = use_ref_end_penalty ? -kOpen - (j - 1) * kExt : 0.0f; tX[j] = 1; Y[j] = -kLarge; M[j] = -kLarge; }
  for (int i = 1; i <= L2; i++) { const size_t c = (size_t)i * W; Y[c] = -kOpen - (i - 1) * kExt; tY[c] = 2; X[c] = -kLarge; M[c] = -kLarge; }
  for (int i = 1; i <= L2; i++)
    for (int j = 1; j <= L1; j++) {
      const size_t h = (size_t)i * W + j, d = h - W - 1, l = h - 1, u = h - W;
      const int a = code(ref[j - 1]), b = code(read[i - 1]);
      M[h] = pick(M[d], X[d], Y[d], &tM[h]) + ((a == 4 || b == 4 || a == b) ? 2.0f : -2.0f);
      X[h] = pick(M[l] - kOpen, X[l] - kExt, Y[l] - kOpen, &tX[h]);
      Y[h] = pick(M[u] - kOpen, X[u] - kOpen, Y[u] - kExt, &tY[h]);
    }
  int col = L1, kind = 0;
  float best;
  const size_t last = (size_t)L2 * W;
  if (use_ref_end_penalty) {
    best = M[last + L1];
    if (X[last + L1] > best) { best = X[last + L1]; kind = 1; }
    if (Y[last + L1] > best) { best = Y[last + L1]; kind = 2; }
  } else {
    best = -kLarge; col = -1; kind = -1;
    for (int j = 0; j <= L1; j++) {
      if (M[last + j] >= best) { best = M[last + j]; col = j; kind = 0; }
      if (X[last + j] > best) { best = X[last + j]; col = j; kind = 1; }
      if (Y[last + j] > best) { best = Y[last + j]; col = j; kind = 2; }
    }
  }
  *score = best;
  std::string out;
  for (int j = L1; j > col; j--) out += 'D';
  int row = L2;
  while (row > 0) {
    const size_t h = (size_t)row * W + col;
    if (kind == 0 && col > 0) { out += 'M'; kind = tM[h]; row--; col--; }
    else if (kind == 1 && col > 0) { out += 'D'; kind = tX[h]; col--; }
    else if (kind == 2) { out += 'I'; kind = tY[h]; row--; }
    else return -1;
  }
  for (; col > 0; col--) out += 'D';
  std::reverse(out.begin(), out.end());
  std::memcpy(ops, out.c_str(), out.size() + 1);
  return (int32_t)out.size();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// SNP phasing log-likelihoods: checker of K7 (hipstr_snp_phasing_batch_host).
// snp_phasing_quality.cpp:4-63 (extract_bases_and_qualities), :65-90 (add_log_phasing_probs),
// :92-120 (calc_het_snp_factors); snp_tree.h:114-126 (findContained = SNPs with start <= pos <= stop,
// in position order).  Two passes per alignment like the reference: first the base / quality under
// every overlapped SNP, then the sums.
// ---------------------------------------------------------------------------------------------
extern "C" int32_t oracle_snp_phasing(const hipstr_snp_phasing_t* b, double* log_p1, double* log_p2, int32_t* counts) {
  for (int e = 0; e < b->n_entries; e++) {
    double p1 = 0.0, p2 = 0.0;
    int32_t c1 = 0, c2 = 0, mis = 0;
    const int set = b->entry_snp_set[e];
    for (int a = b->entry_aln_off[e]; set >= 0 && a < b->entry_aln_off[e + 1]; a++) {
      std::vector<int> snps;   // findContained(Position(), GetEndPosition() - 1)
      const uint32_t* lo = b->snp_pos + b->set_off[set];
      const uint32_t* hi = b->snp_pos + b->set_off[set + 1];
      for (const uint32_t* it = std::lower_bound(lo, hi, (uint32_t)b->aln_pos[a]); it != hi && *it <= (uint32_t)(b->aln_end[a] - 1); ++it)
        snps.push_back((int)(it - b->snp_pos));
      if (snps.empty()) continue;
      const char* seq = b->bases + b->aln_seq_off[a];
      const char* qual = b->quals + b->aln_seq_off[a];
      const int n_bases = b->aln_seq_off[a + 1] - b->aln_seq_off[a];
      std::vector<char> bases, quals;
      int32_t pos = b->aln_pos[a];
      size_t snp_index = 0;
      unsigned int base_index = 0;
      int cigar_index = b->aln_cigar_off[a];
      while (snp_index < snps.size() && cigar_index < b->aln_cigar_off[a + 1]) {
        const uint32_t snp_pos = b->snp_pos[snps[snp_index]];
        const int32_t len = b->cigar_len[cigar_index];
        switch (b->cigar_type[cigar_index]) {
          case 'M': case '=': case 'X':
            if (snp_pos < (uint32_t)(pos + len)) {
              const unsigned int at = snp_pos - pos + base_index;
              if ((int)at >= n_bases) return -2;
              bases.push_back(seq[at]);
              quals.push_back(qual[at]);
              snp_index++;
            } else { pos += len; base_index += len; cigar_index++; }
            break;
          case 'D':
            if (snp_pos < (uint32_t)(pos + len)) { bases.push_back('-'); quals.push_back('-'); snp_index++; }
            else { pos += len; cigar_index++; }
            break;
          case 'I': base_index += len; cigar_index++; break;
          case 'S':
            if (snp_pos < (uint32_t)pos) { bases.push_back('-'); quals.push_back('-'); snp_index++; }
            else { base_index += len; cigar_index++; }
            break;
          case 'H': cigar_index++; break;
          default: return -1;
        }
      }
      if (bases.size() != snps.size()) return -3;   // the reference's assert
      for (size_t i = 0; i < snps.size(); i++) {
        if (bases[i] == '-') continue;
        const unsigned char q = (unsigned char)quals[i];
        if (bases[i] == b->snp_base1[snps[i]]) { p1 += T().qual_correct[q]; p2 += T().qual_error[q]; c1++; }
        else if (bases[i] == b->snp_base2[snps[i]]) { p1 += T().qual_error[q]; p2 += T().qual_correct[q]; c2++; }
        else { p1 += T().qual_error[q]; p2 += T().qual_error[q]; mis++; }
      }
    }
    log_p1[e] = p1; log_p2[e] = p2;
    counts[4 * e] = c1; counts[4 * e + 1] = c2; counts[4 * e + 2] = mis; counts[4 * e + 3] = 0;
  }
  return 0;
}
