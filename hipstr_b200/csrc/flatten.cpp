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
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <thread>
#include <chrono>
#include <functional>
#include <pthread.h>
#include <mutex>
#include <memory>
#include <deque>
#include <condition_variable>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace hipstr {

namespace {
host_alloc_fn g_alloc = nullptr;
host_free_fn g_free = nullptr;
}
void set_host_allocator(host_alloc_fn alloc, host_free_fn release) { g_alloc = alloc; g_free = release; }
void* host_alloc(size_t bytes) {
  void* p = g_alloc ? g_alloc(bytes) : std::malloc(bytes);
  if (!p) throw std::bad_alloc();
  return p;
}
void host_free(void* p) { if (g_free) g_free(p); else std::free(p); }

namespace { thread_local int t_thread_cap = 0; }
void set_host_thread_budget(int n) { t_thread_cap = n > 0 ? n : 0; }
int host_thread_budget() {
  static const int n = [] {
    const char* env = std::getenv("HIPSTR_HOST_THREADS");
    const int t = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(t, 32));
  }();
  return t_thread_cap > 0 ? std::min(n, t_thread_cap) : n;
}

/* ---- the host thread pool ---------------------------------------------------------------------------------------- */
namespace {
struct PoolJob {
  const std::function<void(size_t)>* fn;
  size_t n;
  std::atomic<size_t> next{0}, done{0};
  void work() {
    size_t mine = 0;
    for (size_t i = next.fetch_add(1, std::memory_order_relaxed); i < n; i = next.fetch_add(1, std::memory_order_relaxed)) { (*fn)(i); mine++; }
    if (mine) done.fetch_add(mine, std::memory_order_release);
  }
};
struct ThreadPool {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<PoolJob> > tickets;   // one entry per helper a job asked for
  int n_threads = 0;
  void loop() {
    for (;;) {
      std::shared_ptr<PoolJob> job;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !tickets.empty(); });
        job = std::move(tickets.front());
        tickets.pop_front();
      }
      job->work();   // a ticket taken after the job ran dry finds next >= n and never touches fn
    }
  }
  void grow(int want) {   // under mu
    for (; n_threads < want; n_threads++) std::thread([this] { loop(); }).detach();
  }
};
ThreadPool* g_pool = nullptr;   // leaked on purpose: its parked threads outlive static destruction
std::once_flag g_pool_once;
ThreadPool& thread_pool() {
  std::call_once(g_pool_once, [] {
    g_pool = new ThreadPool();
    // a forked child has none of the parked threads: it starts over with an empty pool
    pthread_atfork(nullptr, nullptr, [] { g_pool = new ThreadPool(); });
  });
  return *g_pool;
}
}  // namespace

void parallel_run(size_t n, int workers, const std::function<void(size_t)>& fn) {
  if (workers > (int)n) workers = (int)n;
  if (workers <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
  auto job = std::make_shared<PoolJob>();
  job->fn = &fn;
  job->n = n;
  ThreadPool& pool = thread_pool();
  {
    std::lock_guard<std::mutex> lk(pool.mu);
    const int hw = std::max(2, std::min((int)std::thread::hardware_concurrency(), 256));
    pool.grow(std::min(hw, std::max(pool.n_threads, (int)pool.tickets.size() + workers - 1)));
    for (int w = 1; w < workers; w++) pool.tickets.push_back(job);
  }
  if (workers == 2) pool.cv.notify_one(); else pool.cv.notify_all();
  job->work();
  // every index is claimed; wait for the helpers still inside fn
  for (int spins = 0; job->done.load(std::memory_order_acquire) < n; spins++) {
    if (spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(20));
    else if (spins > 64) std::this_thread::yield();
  }
}

void FlatBatch::clear() {
  pools.clear(); bases.clear(); quals.clear();
  hapsides.clear(); hapbytes.clear(); blocks.clear(); reps.clear(); progs.clear(); rep_tabs.clear(); hap_mask.clear();
  for (auto& j : jobs) j.clear();
  slot_reps.clear(); locus_slot0.clear(); stut_jobs.clear(); pool_t_off.clear(); chunks.clear();
  stut_n_max = 16;
  n_out = n_alignments = 0;
}

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
  std::vector<uint8_t> codes[2];   // base codes of seq[]
  std::vector<uint8_t> cls0[2];    // homopolymer class of every row ignoring the neighbour blocks
  std::vector<int> sens[2];        // rows whose class can change with the neighbour blocks
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

struct LocusInfo { int32_t hap_rec0, H, max_len; int64_t live_haps; int32_t slot0, n_slots; };
// Bases travel to the device as codes 0..4.  The reference compares raw characters
// (HapAligner.cpp:115,149); restricted to the alphabet ACGTN that is the same relation.
// (A,C,T,G) = (c >> 1) & 3, which the bulk converter below computes without a table; N = 4.
inline int base_code(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'T': return 2;
    case 'G': return 3;
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

// Bulk conversion of read bases to codes, 16 bytes per step (SSE2 is part of x86-64); returns
// non-zero if a byte is outside ACGTN.
inline unsigned convert_bases(const unsigned char* __restrict src, char* __restrict dst, int n) {
  unsigned bad = 0;
  int i = 0;
#if defined(__SSE2__)
  const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T');
  const __m128i cN = _mm_set1_epi8('N'), three = _mm_set1_epi8(3), four = _mm_set1_epi8(4);
  __m128i all_ok = _mm_set1_epi8((char)0xff);
  for (; i + 16 <= n; i += 16) {
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i is_n = _mm_cmpeq_epi8(c, cN);
    const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(c, cA), _mm_cmpeq_epi8(c, cC)),
                                    _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(c, cG), _mm_cmpeq_epi8(c, cT)), is_n));
    all_ok = _mm_and_si128(all_ok, ok);
    const __m128i acgt = _mm_and_si128(_mm_srli_epi16(c, 1), three);
    const __m128i code = _mm_or_si128(_mm_and_si128(is_n, four), _mm_andnot_si128(is_n, acgt));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), code);
  }
  bad |= (_mm_movemask_epi8(all_ok) != 0xffff);
#endif
  for (; i < n; i++) {
    const int x = base_code((char)src[i]);
    bad |= (x < 0);
    dst[i] = (char)x;
  }
  return bad;
}

}  // namespace

int64_t stutter_table_budget_doubles() {
  int64_t mb = 16384;
  if (const char* e = std::getenv("HIPSTR_T_BUDGET_MB")) mb = std::max<int64_t>(1, std::atoll(e));
  return mb * (1 << 20) / 8;
}

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

hipstr_status_t flatten_batch(const hipstr_align_batch_t* b, FlatBatch& out, std::string& err, bool fresh_rows) {
  if (!b || b->n_loci < 0) { err = "null batch"; return HIPSTR_ERR_BAD_ARG; }
  if (b->n_loci > 0 && (!b->locus_block_off || !b->locus_pool_off || !b->locus_hap_off || !b->locus_out_off ||
                        !b->block_period || !b->block_opt_off || !b->block_stutter || !b->opt_seq_off || !b->opt_seq ||
                        !b->pool_seq_off || !b->pool_bases || !b->pool_quals || !b->pool_seed)) {
    err = "null array in batch";
    return HIPSTR_ERR_BAD_ARG;
  }
  out.clear();
  for (int v = 0; v < kNumColVariants; v++) { out.n_max[v] = 16; out.l_max[v] = 2; }
  out.n_out = b->n_loci ? b->locus_out_off[b->n_loci] : 0;

  int64_t total_pairs = count_alignments(b);
  out.n_alignments = total_pairs;
  std::vector<LocusInfo> loci;

  // Per-locus lowering (haplotype sequences, row classes with the reference's reuse history, repeat programs) is the
  // serial bulk of this function; loci are independent, so chunks of loci are lowered by the host threads into private
  // arrays with chunk-local offsets and stitched together afterwards (every cross-reference is an index, rebased once).
  struct Lowered {
    std::vector<DevHapSide> hapsides;
    std::vector<uint8_t> hapbytes, hap_mask;
    std::vector<DevBlock> blocks;
    std::vector<DevRep> reps;
    std::vector<DevProgEntry> progs;
    std::vector<int32_t> rep_tabs;
    std::vector<DevSlotReps> slot_reps;
  };
  auto lower_locus = [b, fresh_rows](int l, Lowered& out, LocusInfo& li_out, std::string& err) -> hipstr_status_t {
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
        if (op.seq[0].size() > 9000) { err = "block option too long"; return HIPSTR_ERR_UNSUPPORTED; }
        op.seq[1].assign(op.seq[0].rbegin(), op.seq[0].rend());
        for (int s = 0; s < 2; s++) {
          op.runs[s].build(op.seq[s]);
          if (!append_codes(op.codes[s], op.seq[s].begin(), op.seq[s].end())) { err = "haplotype bases must be A,C,G,T or N"; return HIPSTR_ERR_UNSUPPORTED; }
          if (blk[k].period == 0) {
            // class of a row = min(15, max(h(c), h(c-1))) with h the homopolymer length around the
            // base (HapAligner.cpp:119-120).  Away from the block ends h only depends on the block.
            const int n = (int)op.seq[s].size();
            const Runs& r = op.runs[s];
            std::vector<int> h0(n);
            std::vector<char> edge(n);
            for (int c = 0; c < n; c++) {
              h0[c] = r.left[c] + r.right[c] + 1;
              edge[c] = (c - r.left[c] == 0) || (c + r.right[c] == n - 1);
            }
            op.cls0[s].resize(n);
            for (int c = 0; c < n; c++) {
              const int cm = std::max(0, c - 1);
              op.cls0[s][c] = (uint8_t)std::min(15, std::max(h0[c], h0[cm]));
              if (edge[c] || edge[cm]) op.sens[s].push_back(c);
            }
          }
        }
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
            out.hapbytes.insert(out.hapbytes.end(), op.codes[s].begin(), op.codes[s].end());
            r.len = B;
            r.period = p;
            r.n_del = n_del;
            r.left_align = (s == 0);
            // upstream_match_lengths_ (StutterAlignerClass.h:35-42,70-75): for lag = k*period,
            // m[i] = 0 if seq[i-lag] != seq[i] else 1 + m[i-1]
            const std::string& sq = op.seq[s];
            const std::vector<uint8_t>& cd = op.codes[s];
            const int n_lag = std::max(n_del, 1);
            std::vector<std::vector<int> > runs(n_lag, std::vector<int>(B, 0));
            for (int k2 = 1; k2 <= n_lag; k2++)
              for (int i = k2 * p; i < B; i++) runs[k2 - 1][i] = sq[i - k2 * p] != sq[i] ? 0 : 1 + runs[k2 - 1][i - 1];
            const int VB = HIPSTR_VAL_STRIDE * 8;   // bytes per read column of the emission table
            auto emit_entry = [&](int pos, int off_a, int off_b, int moves, double logrun) {
              DevProgEntry e;
              e.pos = pos;
              e.moves = moves;
              if (moves) { e.off_a = off_a; e.off_b = off_b; }
              else e.logrun = logrun;
              out.progs.push_back(e);
            };
            const int zero = 0;
            const HostTables& T = host_tables();
            // insertion walk (StutterAlignerClass.cpp:75-97), full extent i > -B; offsets are relative to
            // column j - period, and every further inserted copy is `period` columns further left
            r.prog_off[0] = (int32_t)out.progs.size();
            {
              int i = 0;
              while (i > -B) {
                const int bpos = B - 1 + i;
                int step = 1;
                if (-i + p < B) {
                  const int run = runs[0][bpos];
                  if (run == 0) emit_entry(i, i * VB + cd[bpos] * 8, i * VB + cd[bpos - p] * 8, 1, 0.0);
                  else { emit_entry(i, zero, zero, 0, T.int_logs[run]); step = run; }
                } else
                  emit_entry(i, zero, zero, 0, 0.0);
                i -= step;
              }
              emit_entry(i, zero, zero, 0, 0.0);
            }
            // deletion walks (StutterAlignerClass.cpp:127-142), full extent i > -(B + D); relative to column j
            for (int k2 = 1; k2 <= HIPSTR_MAX_ARTIFACT_UNITS; k2++) {
              r.prog_off[k2] = (int32_t)out.progs.size();
              if (k2 > n_del) continue;
              const int D = -k2 * p;
              int i = 0;
              while (i > -(B + D)) {
                const int bpos = B - 1 + i;
                int step = 1;
                const int run = runs[k2 - 1][bpos];
                if (run == 0) emit_entry(i, i * VB + cd[bpos + D] * 8, i * VB + cd[bpos] * 8, 1, 0.0);
                else { emit_entry(i, zero, zero, 0, T.int_logs[run]); step = run; }
                i -= step;
              }
              emit_entry(i, zero, zero, 0, 0.0);
            }
            // offset tables of the prefix sums of StutterAlignerClass::load_read (:12-53)
            r.diag_off = (int32_t)out.rep_tabs.size();
            for (int t = 0; t < B; t++) out.rep_tabs.push_back(-t * VB + cd[B - 1 - t] * 8);
            r.ins_off = (int32_t)out.rep_tabs.size();
            for (int t = 0; t < HIPSTR_MAX_ARTIFACT_UNITS * p; t++)
              out.rep_tabs.push_back((t % p) < B ? -t * VB + cd[B - 1 - (t % p)] * 8 : -1);
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
    const int slot0 = (int)out.slot_reps.size();
    std::vector<std::vector<int> > slot_of(nb);   // [block][option] -> table slot of the locus, assigned at first use by a live haplotype
    for (int k = 0; k < nb; k++) slot_of[k].assign(n_opts[k], -1);
    std::vector<int32_t> cur(nb), prev(nb);
    std::vector<std::vector<uint8_t> > cached[2];   // [side][oriented block] -> classes of its rows
    cached[0].resize(nb); cached[1].resize(nb);
    std::vector<int> seg1_reps[2];                  // hapside indices, one per distinct seg-1 class
    std::vector<const Option*> opt(nb);
    std::vector<int> period(nb);
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
        for (int k = 0; k < nb; k++) {
          const int src = side == 0 ? k : nb - 1 - k;
          opt[k] = &blk[src].opts[cur[src]];
          period[k] = blk[src].period;
        }
        const int changed = last_changed < 0 ? -1 : (side == 0 ? last_changed : nb - 1 - last_changed);
        hs.seq_off = (int32_t)out.hapbytes.size();
        int len = 0;
        for (int k = 0; k < nb; k++) {
          out.hapbytes.insert(out.hapbytes.end(), opt[k]->codes[side].begin(), opt[k]->codes[side].end());
          len += (int)opt[k]->codes[side].size();
        }
        hs.len = len;
        hs.row_off = (int32_t)out.hapbytes.size();
        out.hapbytes.resize(out.hapbytes.size() + len, HIPSTR_ROW_REPEAT);
        uint8_t* rows = out.hapbytes.data() + hs.row_off;
        hs.blk_off = (int32_t)out.blocks.size();
        hs.n_blocks = nb;
        hs.first_rep = nb;
        int row = 0, first_repeat_row = -1;
        for (int k = 0; k < nb; k++) {
          const int n = (int)opt[k]->seq[side].size();
          DevBlock db = {row, n, period[k] > 0 ? opt[k]->rep[side] : -1, 0};
          if (period[k] > 0) {
            const int src = side == 0 ? k : nb - 1 - k;
            int& slot = slot_of[src][cur[src]];
            if (slot < 0) {
              slot = (int)out.slot_reps.size() - slot0;
              DevSlotReps sr = {opt[k]->rep[side == 0 ? 0 : 1], 0};
              sr.rep_fwd = blk[src].opts[cur[src]].rep[0];
              sr.rep_rev = blk[src].opts[cur[src]].rep[1];
              out.slot_reps.push_back(sr);
            }
            db.tslot = slot;
          }
          out.blocks.push_back(db);
          if (period[k] > 0) {
            if (first_repeat_row < 0) { first_repeat_row = row; hs.first_rep = k; }
          } else {
            hs.n_seed_pos += n;
            std::vector<uint8_t>& keep = cached[side][k];
            if (!(reuse && k < changed)) {   // HapAligner.cpp:54-60: these rows are recomputed now
              keep = opt[k]->cls0[side];
              for (int c : opt[k]->sens[side]) {   // rows whose run reaches a block end see the neighbours
                const int hp = std::max(homopolymer(opt, side, k, c), homopolymer(opt, side, k, std::max(0, c - 1)));
                keep[c] = (uint8_t)std::min(15, hp);
              }
            }
            std::memcpy(rows + row, keep.data(), (size_t)n);
            if (k > 0 && period[k - 1] > 0) rows[row] |= HIPSTR_ROW_AFTER_REPEAT;
          }
          row += n;
        }
        if (first_repeat_row < 0) first_repeat_row = len;
        hs.seg1_class = -1;
        if (nb == 3 && period[0] == 0 && period[1] > 0 && period[2] == 0) {   // the canonical HipSTR haplotype
          const uint8_t* seq = out.hapbytes.data() + hs.seq_off;
          for (size_t c = 0; c < seg1_reps[side].size() && hs.seg1_class < 0; c++) {
            const DevHapSide& o = out.hapsides[seg1_reps[side][c]];
            const int o_first = out.blocks[o.blk_off + 1].row_start;
            if (o_first == first_repeat_row && !std::memcmp(out.hapbytes.data() + o.seq_off, seq, first_repeat_row) &&
                !std::memcmp(out.hapbytes.data() + o.row_off, rows, first_repeat_row))
              hs.seg1_class = (int32_t)c;
          }
          if (hs.seg1_class < 0) {
            hs.seg1_class = (int32_t)seg1_reps[side].size();
            seg1_reps[side].push_back((int)out.hapsides.size());
          }
        }
        max_len = std::max(max_len, len);
        out.hapsides.push_back(hs);
      }
      reuse = live && !fresh_rows;   // HapAligner.cpp:615-619,634: a skipped haplotype breaks the reuse chain
    }
    // DevHapSide offsets into hapbytes were taken while the vector was still growing: they are
    // indices, not pointers, so they stay valid.

    // ---- per-locus facts the pool pass needs ----
    LocusInfo li;
    li.hap_rec0 = hap_rec0;
    li.H = (int32_t)H;
    li.max_len = max_len;
    li.live_haps = H;
    if (mask) { li.live_haps = 0; for (int64_t h = 0; h < H; h++) li.live_haps += mask[h] != 0; }
    li.slot0 = slot0;
    li.n_slots = (int32_t)out.slot_reps.size() - slot0;
    li_out = li;
    return HIPSTR_OK;
  };
  {
    int n_threads = host_thread_budget();
    if (b->n_loci < 64) n_threads = 1;
    const int n_chunks = n_threads == 1 ? 1 : std::min(b->n_loci, n_threads * 4);
    std::vector<Lowered> parts((size_t)n_chunks);
    std::vector<hipstr_status_t> part_status((size_t)n_chunks, HIPSTR_OK);
    std::vector<std::string> part_err((size_t)n_chunks);
    loci.resize((size_t)b->n_loci);
    auto lower_chunk = [&](int c) {
      const int l0 = (int)((int64_t)b->n_loci * c / n_chunks), l1 = (int)((int64_t)b->n_loci * (c + 1) / n_chunks);
      for (int l = l0; l < l1 && part_status[c] == HIPSTR_OK; l++) part_status[c] = lower_locus(l, parts[c], loci[l], part_err[c]);
    };
    if (n_threads == 1) lower_chunk(0);
    else {
      parallel_run((size_t)n_chunks, n_threads, [&](size_t c) { lower_chunk((int)c); });
    }
    for (int c = 0; c < n_chunks; c++)
      if (part_status[c] != HIPSTR_OK) { err = part_err[c]; return part_status[c]; }
    for (int c = 0; c < n_chunks; c++) {
      Lowered& p = parts[c];
      const int32_t side0 = (int32_t)out.hapsides.size(), byte0 = (int32_t)out.hapbytes.size(), blk0 = (int32_t)out.blocks.size(),
                    rep0 = (int32_t)out.reps.size(), prog0 = (int32_t)out.progs.size(), tab0 = (int32_t)out.rep_tabs.size(),
                    slots0 = (int32_t)out.slot_reps.size();
      for (DevHapSide& hs : p.hapsides)
        if (hs.len > 0) { hs.seq_off += byte0; hs.row_off += byte0; hs.blk_off += blk0; }   // (placeholders of masked haplotypes stay zero)
      for (DevBlock& db : p.blocks)
        if (db.rep >= 0) db.rep += rep0;
      for (DevRep& r : p.reps) {
        r.seq_off += byte0;
        for (int k = 0; k <= HIPSTR_MAX_ARTIFACT_UNITS; k++) r.prog_off[k] += prog0;
        r.diag_off += tab0;
        r.ins_off += tab0;
      }
      const int l0 = (int)((int64_t)b->n_loci * c / n_chunks), l1 = (int)((int64_t)b->n_loci * (c + 1) / n_chunks);
      for (int l = l0; l < l1; l++) { loci[l].hap_rec0 += side0; loci[l].slot0 += slots0; }
      for (DevSlotReps& sr : p.slot_reps) { sr.rep_fwd += rep0; sr.rep_rev += rep0; }
      out.slot_reps.insert(out.slot_reps.end(), p.slot_reps.begin(), p.slot_reps.end());
      out.hapsides.insert(out.hapsides.end(), p.hapsides.begin(), p.hapsides.end());
      out.hapbytes.insert(out.hapbytes.end(), p.hapbytes.begin(), p.hapbytes.end());
      out.hap_mask.insert(out.hap_mask.end(), p.hap_mask.begin(), p.hap_mask.end());
      out.blocks.insert(out.blocks.end(), p.blocks.begin(), p.blocks.end());
      out.reps.insert(out.reps.end(), p.reps.begin(), p.reps.end());
      out.progs.insert(out.progs.end(), p.progs.begin(), p.progs.end());
      out.rep_tabs.insert(out.rep_tabs.end(), p.rep_tabs.begin(), p.rep_tabs.end());
      p = Lowered();
    }
  }

  // the kernel prefetches two program entries ahead: pad the arrays
  for (int k = 0; k < 2; k++) {
    DevProgEntry e;
    e.pos = -(1 << 30); e.moves = 0; e.logrun = 0.0;
    out.progs.push_back(e);
  }

  // ---- pooled reads: offsets (serial prefix), then a threaded fill of the big byte arrays ----
  const int n_pools = b->n_pools;
  std::vector<int32_t> pool_off((size_t)n_pools + 1);
  {
    int64_t at = 0;
    for (int p = 0; p < n_pools; p++) {
      const int n = b->pool_seq_off[p + 1] - b->pool_seq_off[p];
      if (n < 0) { err = "pool_seq_off not monotone"; return HIPSTR_ERR_BAD_ARG; }
      pool_off[p] = (int32_t)at;
      at += round_up(std::max(n, 1), 16);
      if (at > 0x7fffffff) { err = "more than 2 GiB of read bases in one batch"; return HIPSTR_ERR_UNSUPPORTED; }
    }
    pool_off[n_pools] = (int32_t)at;
    out.bases.resize((size_t)at);
    out.quals.resize((size_t)at);
    out.pools.resize((size_t)n_pools);
  }
  std::vector<uint8_t> pool_variant((size_t)n_pools, 255);   // 255 = no job, 254 = seedless
  std::atomic<int> status(HIPSTR_OK);
  auto fill = [&](int l0, int l1) {
    for (int l = l0; l < l1 && status.load(std::memory_order_relaxed) == HIPSTR_OK; l++) {
      const LocusInfo& li = loci[l];
      for (int p = b->locus_pool_off[l]; p < b->locus_pool_off[l + 1]; p++) {
        const int s0 = b->pool_seq_off[p], n = b->pool_seq_off[p + 1] - s0;
        const int padded = pool_off[p + 1] - pool_off[p];
        DevPool dp;
        dp.seq_off = pool_off[p];
        dp.len = n;
        dp.seed = b->pool_seed[p];
        dp.locus = l;
        dp.out_off = b->locus_out_off[l] + (int64_t)(p - b->locus_pool_off[l]) * li.H;
        dp.hap_rec0 = li.hap_rec0;
        dp.n_haps = li.H;
        out.pools[p] = dp;
        char* bd = out.bases.data() + pool_off[p];
        char* qd = out.quals.data() + pool_off[p];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(b->pool_bases + s0);
        const unsigned bad = convert_bases(src, bd, n);
        std::memcpy(qd, b->pool_quals + s0, (size_t)n);
        for (int i = n; i < padded; i++) { bd[i] = 4; qd[i] = '!'; }
        if (bad) { status.store(HIPSTR_ERR_UNSUPPORTED); return; }   // a base outside ACGTN
        if (b->realign_pool && !b->realign_pool[p]) continue;
        if (dp.seed < 0) { pool_variant[p] = 254; continue; }
        if (dp.seed == 0 || dp.seed >= n - 1) { status.store(HIPSTR_ERR_INVALID_SEED); return; }
        if (li.live_haps == 0) continue;
        const int v = pick_variant(dp.seed, n - dp.seed - 1);
        if (v < 0) { status.store(HIPSTR_ERR_BAD_ARG + 100); return; }
        pool_variant[p] = (uint8_t)v;
      }
    }
  };
  int n_threads = host_thread_budget();
  if (n_pools < 20000) n_threads = 1;
  if (n_threads == 1) fill(0, b->n_loci);
  else {
    std::vector<std::pair<int, int> > ranges;
    int l0 = 0;
    for (int t = 0; t < n_threads; t++) {   // split by pool count, not locus count
      const int64_t target = (int64_t)n_pools * (t + 1) / n_threads;
      int l1 = l0;
      while (l1 < b->n_loci && b->locus_pool_off[l1 + 1] <= target) l1++;
      if (t == n_threads - 1) l1 = b->n_loci;
      if (l1 > l0) ranges.emplace_back(l0, l1);
      l0 = l1;
    }
    parallel_run(ranges.size(), n_threads, [&](size_t i) { fill(ranges[i].first, ranges[i].second); });
  }
  switch (status.load()) {
    case HIPSTR_OK: break;
    case HIPSTR_ERR_UNSUPPORTED: err = "read bases must be A,C,G,T or N"; return HIPSTR_ERR_UNSUPPORTED;
    case HIPSTR_ERR_INVALID_SEED: err = "invalid alignment seed"; return HIPSTR_ERR_INVALID_SEED;
    default: err = "read longer than the kernel's limit"; return HIPSTR_ERR_UNSUPPORTED;
  }

  out.locus_slot0.resize((size_t)b->n_loci);
  for (int l = 0; l < b->n_loci; l++) out.locus_slot0[l] = loci[l].slot0;

  // ---- jobs ----
  // Split a pool's haplotypes over several warps only when the batch is too small to fill the GPU.
  // Jobs are emitted in pool order, so a chunk of pools (stutter tables within the budget) is a contiguous range of
  // every job list.
  const int64_t budget = stutter_table_budget_doubles();
  out.pool_t_off.resize((size_t)n_pools);
  size_t count[kNumColVariants] = {0}, n_stut = 0;
  for (int l = 0; l < b->n_loci; l++) {
    const LocusInfo& li = loci[l];
    int chunk = li.H;
    if (total_pairs > 0 && total_pairs / li.H < 16384) chunk = (int)std::max<int64_t>(1, std::min<int64_t>(li.H, total_pairs / 16384));
    const int per_pool = (li.H + chunk - 1) / chunk;
    for (int p = b->locus_pool_off[l]; p < b->locus_pool_off[l + 1]; p++) {
      const uint8_t v = pool_variant[p];
      if (v == 255) continue;
      if (v == 254) { count[0]++; continue; }
      count[v] += per_pool;
      n_stut += (size_t)(li.n_slots + HIPSTR_STUT_SLOTS_PER_JOB - 1) / HIPSTR_STUT_SLOTS_PER_JOB;
      const int len = b->pool_seq_off[p + 1] - b->pool_seq_off[p];
      out.n_max[v] = std::max(out.n_max[v], round_up(len, 4));
      out.l_max[v] = std::max(out.l_max[v], round_up(li.max_len, 2));
      out.stut_n_max = std::max(out.stut_n_max, round_up(len, 16));
    }
  }
  size_t at[kNumColVariants], at_stut = 0;
  for (int v = 0; v < kNumColVariants; v++) { out.jobs[v].resize(count[v]); at[v] = 0; }
  out.stut_jobs.resize(n_stut);
  FlatBatch::Chunk ck;
  auto open_chunk = [&] {
    ck.stut_job0 = (int32_t)at_stut;
    for (int v = 0; v < kNumColVariants; v++) ck.job0[v] = (int32_t)at[v];
    ck.t_doubles = 0;
  };
  auto close_chunk = [&] {
    ck.stut_job1 = (int32_t)at_stut;
    bool any = ck.stut_job1 > ck.stut_job0;
    for (int v = 0; v < kNumColVariants; v++) { ck.job1[v] = (int32_t)at[v]; any |= ck.job1[v] > ck.job0[v]; }
    if (any) out.chunks.push_back(ck);
  };
  open_chunk();
  for (int l = 0; l < b->n_loci; l++) {
    const LocusInfo& li = loci[l];
    int chunk = li.H;
    if (total_pairs > 0 && total_pairs / li.H < 16384) chunk = (int)std::max<int64_t>(1, std::min<int64_t>(li.H, total_pairs / 16384));
    for (int p = b->locus_pool_off[l]; p < b->locus_pool_off[l + 1]; p++) {
      const uint8_t v = pool_variant[p];
      out.pool_t_off[p] = 0;
      if (v == 255) continue;
      if (v == 254) {   // HapAligner.cpp:333-337: LL 0 for every haplotype, mask ignored
        DevJob j = {p, 0, li.H, 0};
        out.jobs[0][at[0]++] = j;
        continue;
      }
      const int len = b->pool_seq_off[p + 1] - b->pool_seq_off[p];
      const int64_t need = (int64_t)li.n_slots * HIPSTR_NUM_ARTIFACTS * hipstr_t_pitch(len);
      if (ck.t_doubles > 0 && ck.t_doubles + need > budget) { close_chunk(); open_chunk(); }
      out.pool_t_off[p] = ck.t_doubles;
      ck.t_doubles += need;
      for (int s0 = 0; s0 < li.n_slots; s0 += HIPSTR_STUT_SLOTS_PER_JOB) {
        DevStutJob sj = {p, li.slot0 + s0, std::min(HIPSTR_STUT_SLOTS_PER_JOB, li.n_slots - s0), s0};
        out.stut_jobs[at_stut++] = sj;
      }
      for (int h0 = 0; h0 < li.H; h0 += chunk) {
        DevJob j = {p, h0, std::min(li.H, h0 + chunk), 0};
        out.jobs[v][at[v]++] = j;
      }
    }
  }
  close_chunk();
  return HIPSTR_OK;
}

}  // namespace hipstr
