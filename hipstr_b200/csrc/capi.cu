/*
 * capi.cu -- the C-ABI of include/hipstr_b200.h: context, host<->device staging, launches.
 * No CPU fallback: every compute entry point needs a context, and a context needs a GPU.
 */
#include <chrono>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "flatten.h"
#include "kernels.h"
#include "layout.h"

using namespace hipstr;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct hipstr_dev_batch {
  DevBuf pools, bases, quals, hapsides, hapbytes, blocks, reps, progs, tabs, mask, jobs[kNumColVariants];
  DevBuf slot_reps, stut_jobs, pool_t_off;      // K1a: stutter-table slots, jobs, slab offsets
  std::vector<FlatBatch::Chunk> chunks;         // pool ranges whose stutter tables fit the table buffer
  int32_t stut_n_max = 16;
  int32_t n_jobs[kNumColVariants];
  int32_t n_max[kNumColVariants], l_max[kNumColVariants];
  int64_t n_out = 0, n_alignments = 0;
  bool has_mask = false;
  void release() {
    pools.release(); bases.release(); quals.release(); hapsides.release(); hapbytes.release();
    blocks.release(); reps.release(); progs.release(); tabs.release(); mask.release();
    slot_reps.release(); stut_jobs.release(); pool_t_off.release();
    for (auto& j : jobs) j.release();
  }
};

struct hipstr_dev_genotype {
  hipstr_dev_batch align;
  DevBuf loci, samples, locus_sample_off, pool_seed, pool_index, second_mate, copy_read, read_weight, log_p1, log_p2;
  DevBuf pool_ll;                      // K1 output, [n_out]
  int32_t n_loci = 0, n_samples = 0, n_reads = 0;
  int64_t n_elems = 0, post_size = 0;
  bool has_copy_read = false, masked = false;
  void release() {
    align.release();
    loci.release(); samples.release(); locus_sample_off.release(); pool_seed.release(); pool_index.release();
    second_mate.release(); copy_read.release(); read_weight.release(); log_p1.release(); log_p2.release();
    pool_ll.release();
  }
};

struct StageEvents { cudaEvent_t e[3]; };

struct hipstr_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  double *d_qual_lut = nullptr, *d_trans = nullptr, *d_int_logs = nullptr;
  std::string last_error;
  int32_t last_launches = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  bool timing = false;
  float last_ms = 0.f;
  std::vector<StageEvents> pending;      // recorded, not yet read
  std::vector<StageEvents> free_events;
  FlatBatch flat;                        // host staging (page-locked), reused by every call
  hipstr_dev_batch scratch;              // reused by hipstr_align_batch_host
  hipstr_dev_genotype gscratch;          // reused by hipstr_genotype_batch_host
  DevBuf d_ll, d_pos, d_misc[12], d_out[6], d_last, d_counters;
  DevBuf d_stut, d_stut2;                // stutter tables of the chunks in flight (K1a -> K1b), two alternating buffers
  cudaStream_t table_stream = nullptr;   // K1a of chunk c+1 runs here while K1b of chunk c drains on `stream`
  cudaEvent_t ev_tables[2] = {nullptr, nullptr}, ev_folded[2] = {nullptr, nullptr}, ev_inputs = nullptr;
  cudaEvent_t ev_wait = nullptr;         // cudaEventBlockingSync: host waits sleep instead of spinning (see wait_stream)
  bool sleeping_waits = true;
  DevBuf d_stut_pos, d_dec, d_art, d_job_t_off[kNumColVariants];   // K5 forward pass -> walk back
  double* d_debug = nullptr;             // test hook, see hipstr_debug_lastcols
  double trace_seconds[4] = {0, 0, 0, 0};   // accumulated over hipstr_trace_batch_host calls: lowering, ordering + uploads, kernel, downloads
};

namespace {

hipstr_status_t fail(hipstr_ctx* c, hipstr_status_t st, const std::string& msg) {
  if (c) c->last_error = msg;
  return st;
}

// Host wait for everything queued on `s`.  cudaStreamSynchronize spins on a core; the loop runs several pipelines per GPU and
// several GPUs per box on the same cores, where a spinning wait takes a core away from another pipeline's host work, so the
// wait goes through an event created with cudaEventBlockingSync (the thread sleeps) -- after polling for 150 us first, which
// keeps the latency of small calls (a few traces of one locus) where it was.  HIPSTR_SPIN_WAITS=1 restores spinning.
cudaError_t wait_stream(hipstr_ctx* ctx, cudaStream_t s) {
  if (!ctx->sleeping_waits || !ctx->ev_wait) return cudaStreamSynchronize(s);
  cudaError_t e = cudaEventRecord(ctx->ev_wait, s);
  if (e != cudaSuccess) return e;
  const auto t0 = std::chrono::steady_clock::now();
  while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150)) {
    e = cudaEventQuery(ctx->ev_wait);
    if (e != cudaErrorNotReady) return e;   // done (cudaSuccess) or failed
  }
  return cudaEventSynchronize(ctx->ev_wait);
}
#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(ctx, HIPSTR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

thread_local int64_t g_h2d = 0;   // bytes queued host->device by put() since begin_call()
template <class T>
cudaError_t put(DevBuf& d, const T* v, size_t n, cudaStream_t s) {
  cudaError_t e = d.reserve(std::max<size_t>(n * sizeof(T), 16));
  if (e != cudaSuccess || n == 0) return e;
  g_h2d += (int64_t)(n * sizeof(T));
  return cudaMemcpyAsync(d.p, v, n * sizeof(T), cudaMemcpyHostToDevice, s);
}
template <class T>
cudaError_t put(DevBuf& d, const std::vector<T>& v, cudaStream_t s) { return put(d, v.data(), v.size(), s); }
template <class T>
cudaError_t put(DevBuf& d, const HostBuf<T>& v, cudaStream_t s) { return put(d, v.data(), v.size(), s); }

// Page-locked staging for the big host arrays (flatten.cpp's packed batch, the loop's packing / result buffers):
// H2D / D2H copies then run at full PCIe speed and truly asynchronously.  cudaHostAlloc costs milliseconds, and the loop
// asks for the same few sizes every round, so freed blocks are kept in a size-sorted cache and handed out again
// (best fit, at most twice the request); the cache is capped, the excess goes back to the driver.
struct PinnedCache {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> size_of;
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCached = (size_t)8 << 30;
};
PinnedCache& pinned_cache() { static PinnedCache* c = new PinnedCache(); return *c; }
void* pinned_alloc(size_t bytes) {
  PinnedCache& c = pinned_cache();
  if (bytes == 0) bytes = 16;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_blocks.lower_bound(bytes);
    if (it != c.free_blocks.end() && it->first <= 2 * bytes + 4096) {
      void* p = it->second;
      c.cached_bytes -= it->first;
      c.free_blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(c.mu);
  c.size_of[p] = bytes;
  return p;
}
void pinned_free(void* p) {
  if (!p) return;
  PinnedCache& c = pinned_cache();
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.size_of.find(p);
    if (it == c.size_of.end()) { std::free(p); return; }   // allocated with malloc before a context existed
    bytes = it->second;
    if (c.cached_bytes + bytes <= PinnedCache::kMaxCached) {
      c.free_blocks.emplace(bytes, p);
      c.cached_bytes += bytes;
      return;
    }
    c.size_of.erase(it);
  }
  cudaFreeHost(p);
}
void begin_call(hipstr_ctx* c) { g_h2d = 0; c->h2d_bytes = c->d2h_bytes = 0; c->last_launches = 0; }
void end_call(hipstr_ctx* c) { c->h2d_bytes = g_h2d; }
template <class T>
cudaError_t get(hipstr_ctx* c, T* dst, const void* src, size_t n) {
  if (n == 0) return cudaSuccess;
  c->d2h_bytes += (int64_t)(n * sizeof(T));
  return cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream);
}

hipstr_status_t stage(hipstr_ctx* ctx, const hipstr_align_batch_t* batch, hipstr_dev_batch& d) {
  FlatBatch& f = ctx->flat;
  std::string err;
  hipstr_status_t st;
  try {
    st = flatten_batch(batch, f, err);
  } catch (const std::bad_alloc&) {
    return fail(ctx, HIPSTR_ERR_CUDA, "out of (page-locked) host memory while staging the batch");
  }
  if (st != HIPSTR_OK) return fail(ctx, st, err);
  cudaStream_t s = ctx->stream;
  CU(put(d.pools, f.pools, s));
  CU(put(d.bases, f.bases, s));
  CU(put(d.quals, f.quals, s));
  CU(put(d.hapsides, f.hapsides, s));
  CU(put(d.hapbytes, f.hapbytes, s));
  CU(put(d.blocks, f.blocks, s));
  CU(put(d.reps, f.reps, s));
  CU(put(d.progs, f.progs, s));
  CU(put(d.tabs, f.rep_tabs, s));
  CU(put(d.mask, f.hap_mask, s));
  CU(put(d.slot_reps, f.slot_reps, s));
  CU(put(d.stut_jobs, f.stut_jobs, s));
  CU(put(d.pool_t_off, f.pool_t_off, s));
  d.chunks = f.chunks;
  d.stut_n_max = f.stut_n_max;
  for (int v = 0; v < kNumColVariants; v++) {
    CU(put(d.jobs[v], f.jobs[v], s));
    d.n_jobs[v] = (int32_t)f.jobs[v].size();
    d.n_max[v] = f.n_max[v];
    d.l_max[v] = f.l_max[v];
  }
  d.n_out = f.n_out;
  d.n_alignments = f.n_alignments;
  d.has_mask = !f.hap_mask.empty();
  // the staging buffers are reused by the next call: the copies above must have completed
  CU(wait_stream(ctx, s));
  return HIPSTR_OK;
}

hipstr_status_t run_align(hipstr_ctx* ctx, const hipstr_dev_batch& d, double* ll_dev, int32_t* pos_dev) {
  AlignParams p;
  std::memset(&p, 0, sizeof(p));
  p.pools = (const DevPool*)d.pools.p;
  p.bases = (const char*)d.bases.p;
  p.quals = (const char*)d.quals.p;
  p.hapsides = (const DevHapSide*)d.hapsides.p;
  p.hapbytes = (const uint8_t*)d.hapbytes.p;
  p.blocks = (const DevBlock*)d.blocks.p;
  p.reps = (const DevRep*)d.reps.p;
  p.progs = (const DevProgEntry*)d.progs.p;
  p.rep_tabs = (const int32_t*)d.tabs.p;
  p.hap_mask = d.has_mask ? (const uint8_t*)d.mask.p : nullptr;
  p.qual_lut = ctx->d_qual_lut;
  p.trans = ctx->d_trans;
  p.int_logs = ctx->d_int_logs;
  p.ll_out = ll_dev;
  p.pos_out = pos_dev;
  int l_all = 2;
  for (int v = 0; v < kNumColVariants; v++) if (d.n_jobs[v]) l_all = std::max(l_all, d.l_max[v]);
  CU(ctx->d_last.reserve((size_t)HIPSTR_MAX_ALIGN_CTAS * HIPSTR_WARPS_PER_CTA * 2 * l_all * sizeof(double)));
  // one job counter per launch: K1a + up to kNumColVariants K1b launches per chunk
  const size_t n_counters = std::max<size_t>(1, d.chunks.size()) * (kNumColVariants + 1);
  CU(ctx->d_counters.reserve(n_counters * sizeof(int32_t)));
  CU(cudaMemsetAsync(ctx->d_counters.p, 0, n_counters * sizeof(int32_t), ctx->stream));
  int64_t t_max = 2;
  for (const FlatBatch::Chunk& ck : d.chunks) t_max = std::max(t_max, ck.t_doubles);
  CU(ctx->d_stut.reserve((size_t)t_max * sizeof(double)));
  // With more than one chunk the tables alternate between two buffers and K1a runs on its own stream: the tables of
  // chunk c+1 are computed while K1b of chunk c drains, so no launch ends on a GPU that is emptying.
  const bool overlap = d.chunks.size() > 1;
  if (overlap) CU(ctx->d_stut2.reserve((size_t)t_max * sizeof(double)));
  double* const tables[2] = {(double*)ctx->d_stut.p, overlap ? (double*)ctx->d_stut2.p : (double*)ctx->d_stut.p};
  p.last_scratch = (double*)ctx->d_last.p;
  p.pool_t_off = (const int64_t*)d.pool_t_off.p;
  p.l_max = l_all;
  p.debug_out = ctx->d_debug;
  StutParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.n_max = d.stut_n_max;
  sp.pools = p.pools; sp.bases = p.bases; sp.quals = p.quals;
  sp.slot_reps = (const DevSlotReps*)d.slot_reps.p;
  sp.reps = p.reps; sp.progs = p.progs; sp.rep_tabs = p.rep_tabs;
  sp.qual_lut = p.qual_lut; sp.int_logs = p.int_logs;
  sp.pool_t_off = p.pool_t_off;
  cudaStream_t table_stream = overlap ? ctx->table_stream : ctx->stream;
  if (overlap) {   // the uploads and the counter reset above are on `stream`
    CU(cudaEventRecord(ctx->ev_inputs, ctx->stream));
    CU(cudaStreamWaitEvent(table_stream, ctx->ev_inputs, 0));
  }
  for (size_t c = 0; c < d.chunks.size(); c++) {
    const FlatBatch::Chunk& ck = d.chunks[c];
    const int buf = (int)(c & 1);
    int32_t* counters = (int32_t*)ctx->d_counters.p + c * (kNumColVariants + 1);
    // K1a: the stutter tables of the chunk's (read, allele) pairs, once K1b of the chunk that used this buffer is done
    if (overlap && c >= 2) CU(cudaStreamWaitEvent(table_stream, ctx->ev_folded[buf], 0));
    if (ck.stut_job1 > ck.stut_job0) {
      sp.jobs = (const DevStutJob*)d.stut_jobs.p + ck.stut_job0;
      sp.n_jobs = ck.stut_job1 - ck.stut_job0;
      sp.job_counter = counters + kNumColVariants;
      sp.stut = tables[buf];
      CU(launch_stutter(sp, table_stream));
      ctx->last_launches++;
    }
    if (overlap) {
      CU(cudaEventRecord(ctx->ev_tables[buf], table_stream));
      CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_tables[buf], 0));
    }
    p.stut = tables[buf];
    for (int v = kNumColVariants - 1; v >= 0; v--) {   // longest reads first
      if (ck.job1[v] == ck.job0[v]) continue;
      p.jobs = (const DevJob*)d.jobs[v].p + ck.job0[v];
      p.n_jobs = ck.job1[v] - ck.job0[v];
      p.n_max = d.n_max[v];
      p.job_counter = counters + v;
      CU(launch_align(v, p, HIPSTR_MAX_ALIGN_CTAS, ctx->stream, nullptr));
      ctx->last_launches++;
    }
    if (overlap) CU(cudaEventRecord(ctx->ev_folded[buf], ctx->stream));
  }
  return HIPSTR_OK;
}

// stage timing: events are recorded on the stream and read later (hipstr_collect_timing), so
// enabling it does not serialise the steps
hipstr_status_t mark(hipstr_ctx* ctx, StageEvents& ev, int which) {
  if (!ctx->timing) return HIPSTR_OK;
  if (which == 0) {
    if (ctx->free_events.empty()) {
      for (auto& e : ev.e) CU(cudaEventCreate(&e));
    } else {
      ev = ctx->free_events.back();
      ctx->free_events.pop_back();
    }
  }
  CU(cudaEventRecord(ev.e[which], ctx->stream));
  if (which == 2) ctx->pending.push_back(ev);
  return HIPSTR_OK;
}

hipstr_status_t stage_reads(hipstr_ctx* ctx, const hipstr_align_batch_t* b, const hipstr_reads_batch_t* r,
                            hipstr_dev_genotype& g) {
  const int n_loci = b->n_loci;
  if (n_loci > 0 && (!r->locus_read_off || !r->locus_sample_off || !r->pool_index || !r->sample_label ||
                     !r->second_mate || !r->read_weight || !r->log_p1 || !r->log_p2 || !r->haploid))
    return fail(ctx, HIPSTR_ERR_BAD_ARG, "null array in reads batch");
  const int32_t R = n_loci ? r->locus_read_off[n_loci] : 0, S = n_loci ? r->locus_sample_off[n_loci] : 0;
  std::vector<ScatterLocus> loci((size_t)n_loci);
  std::vector<PostSample> samples((size_t)S);
  int64_t elem = 0, post = 0, hap0 = 0;
  for (int l = 0; l < n_loci; l++) {
    const int64_t H64 = b->locus_hap_off[l + 1] - b->locus_hap_off[l];
    if (H64 <= 0 || H64 + 1 >= 10000) return fail(ctx, HIPSTR_ERR_UNSUPPORTED, "haplotype count outside 1..9998");
    const int H = (int)H64;
    const int r0 = r->locus_read_off[l], r1 = r->locus_read_off[l + 1];
    const int P = b->locus_pool_off[l + 1] - b->locus_pool_off[l];
    ScatterLocus& L = loci[l];
    L.elem_off = elem;
    L.pool_ll_off = b->locus_out_off[l];
    L.read0 = r0;
    L.n_reads = r1 - r0;
    L.n_haps = H;
    L.pool0 = b->locus_pool_off[l];
    L.hap0 = hap0;
    for (int i = r0; i < r1; i++)
      if (r->pool_index[i] < 0 || r->pool_index[i] >= P) return fail(ctx, HIPSTR_ERR_BAD_ARG, "pool_index out of range");
    const int s0 = r->locus_sample_off[l], s1 = r->locus_sample_off[l + 1];
    int i = r0;
    for (int s = s0; s < s1; s++) {
      PostSample& ps = samples[s];
      ps.read0 = i;
      while (i < r1 && r->sample_label[i] == s - s0) i++;   // reads are sample-major (genotyper.h:104-112)
      ps.read1 = i;
      ps.n_haps = H;
      ps.haploid = r->haploid[l];
      ps.ll_off = elem;
      ps.locus_read0 = r0;
      ps.locus = l;
      ps.post_off = post + (int64_t)(s - s0) * H * H;
    }
    if (i != r1) return fail(ctx, HIPSTR_ERR_BAD_ARG, "reads are not sample-major or a sample label is out of range");
    elem += (int64_t)(r1 - r0) * H;
    post += (int64_t)(s1 - s0) * H * H;
    hap0 += H;
  }
  cudaStream_t s = ctx->stream;
  CU(put(g.loci, loci, s));
  CU(put(g.samples, samples, s));
  CU(put(g.locus_sample_off, r->locus_sample_off, (size_t)n_loci + 1, s));
  CU(put(g.pool_seed, b->pool_seed, (size_t)b->n_pools, s));
  CU(put(g.pool_index, r->pool_index, (size_t)R, s));
  CU(put(g.second_mate, r->second_mate, (size_t)R, s));
  if (r->copy_read) CU(put(g.copy_read, r->copy_read, (size_t)R, s));
  CU(put(g.read_weight, r->read_weight, (size_t)R, s));
  CU(put(g.log_p1, r->log_p1, (size_t)R, s));
  CU(put(g.log_p2, r->log_p2, (size_t)R, s));
  CU(wait_stream(ctx, s));   // `loci` / `samples` are locals
  g.n_loci = n_loci; g.n_samples = S; g.n_reads = R; g.n_elems = elem; g.post_size = post;
  g.has_copy_read = r->copy_read != nullptr;
  g.masked = r->copy_read || b->realign_pool || b->realign_hap;
  return HIPSTR_OK;
}

hipstr_status_t run_genotype(hipstr_ctx* ctx, hipstr_dev_genotype& g, const hipstr_genotype_out_t& o) {
  StageEvents ev;
  hipstr_status_t st;
  CU(g.pool_ll.reserve(std::max<size_t>((size_t)g.align.n_out * sizeof(double), 16)));
  if ((st = mark(ctx, ev, 0)) != HIPSTR_OK) return st;
  if ((st = run_align(ctx, g.align, (double*)g.pool_ll.p, nullptr)) != HIPSTR_OK) return st;
  if ((st = mark(ctx, ev, 1)) != HIPSTR_OK) return st;
  ScatterBatchParams sp;
  sp.n_loci = g.n_loci; sp.n_elems = g.n_elems;
  sp.loci = (const ScatterLocus*)g.loci.p;
  sp.pool_ll = (const double*)g.pool_ll.p;
  sp.pool_seed = (const int32_t*)g.pool_seed.p;
  sp.pool_index = (const int32_t*)g.pool_index.p;
  sp.second_mate = (const uint8_t*)g.second_mate.p;
  sp.copy_read = g.has_copy_read ? (const uint8_t*)g.copy_read.p : nullptr;
  sp.hap_mask = g.align.has_mask ? (const uint8_t*)g.align.mask.p : nullptr;
  sp.read_ll = o.read_ll;
  sp.read_seed = o.read_seed;
  CU(launch_scatter_batch(sp, ctx->stream));
  if (g.n_elems > 0) ctx->last_launches++;
  PostParams pp;
  pp.n_samples = g.n_samples; pp.n_loci = g.n_loci;
  pp.samples = (const PostSample*)g.samples.p;
  pp.locus_sample_off = (const int32_t*)g.locus_sample_off.p;
  pp.read_ll = o.read_ll;
  pp.log_p1 = (const double*)g.log_p1.p; pp.log_p2 = (const double*)g.log_p2.p;
  pp.read_weight = (const int32_t*)g.read_weight.p;
  pp.int_logs = ctx->d_int_logs; pp.log_one_half = host_tables().log_one_half;
  pp.post_out = o.post; pp.sample_ll_out = o.sample_ll; pp.best_out = o.best; pp.total_ll_out = o.total_ll;
  CU(launch_posteriors(pp, ctx->stream));
  if (g.n_samples > 0) ctx->last_launches += o.total_ll ? 2 : 1;
  return mark(ctx, ev, 2);
}

}  // namespace

extern "C" {

const char* hipstr_version(void) { return "hipstr_b200 0.1.0 (sm_100a)"; }

hipstr_status_t hipstr_create(int device, hipstr_ctx_t** out_ctx) {
  if (!out_ctx) return HIPSTR_ERR_BAD_ARG;
  *out_ctx = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return HIPSTR_ERR_NO_DEVICE;
  }
  set_host_allocator(pinned_alloc, pinned_free);
  hipstr_ctx* ctx = new hipstr_ctx();
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t e) {
    std::fprintf(stderr, "hipstr_create: %s: %s\n", what, cudaGetErrorString(e));
    hipstr_destroy(ctx);
    return HIPSTR_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaStreamCreateWithFlags(&ctx->table_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  for (cudaEvent_t* ev : {&ctx->ev_tables[0], &ctx->ev_tables[1], &ctx->ev_folded[0], &ctx->ev_folded[1], &ctx->ev_inputs})
    if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_wait, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) return bail("event", e);
  if (const char* spin = std::getenv("HIPSTR_SPIN_WAITS")) ctx->sleeping_waits = std::atoi(spin) == 0;
  const HostTables& T = host_tables();
  if ((e = cudaMalloc(&ctx->d_qual_lut, sizeof(T.qual_lut))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->d_trans, sizeof(T.trans))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->d_int_logs, sizeof(T.int_logs))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpy(ctx->d_qual_lut, T.qual_lut, sizeof(T.qual_lut), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", e);
  if ((e = cudaMemcpy(ctx->d_trans, T.trans, sizeof(T.trans), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", e);
  if ((e = cudaMemcpy(ctx->d_int_logs, T.int_logs, sizeof(T.int_logs), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", e);
  *out_ctx = ctx;
  return HIPSTR_OK;
}

void hipstr_destroy(hipstr_ctx_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  if (ctx->table_stream) { cudaStreamSynchronize(ctx->table_stream); cudaStreamDestroy(ctx->table_stream); }
  for (cudaEvent_t ev : {ctx->ev_tables[0], ctx->ev_tables[1], ctx->ev_folded[0], ctx->ev_folded[1], ctx->ev_inputs, ctx->ev_wait})
    if (ev) cudaEventDestroy(ev);
  ctx->d_stut2.release();
  ctx->scratch.release();
  ctx->gscratch.release();
  ctx->d_ll.release();
  ctx->d_pos.release();
  ctx->d_last.release();
  ctx->d_counters.release();
  ctx->d_stut.release(); ctx->d_stut_pos.release(); ctx->d_dec.release(); ctx->d_art.release();
  for (auto& b : ctx->d_job_t_off) b.release();
  for (auto& b : ctx->d_misc) b.release();
  for (auto& b : ctx->d_out) b.release();
  for (auto& ev : ctx->pending) for (auto& e : ev.e) cudaEventDestroy(e);
  for (auto& ev : ctx->free_events) for (auto& e : ev.e) cudaEventDestroy(e);
  if (ctx->d_debug) cudaFree(ctx->d_debug);
  if (ctx->d_qual_lut) cudaFree(ctx->d_qual_lut);
  if (ctx->d_trans) cudaFree(ctx->d_trans);
  if (ctx->d_int_logs) cudaFree(ctx->d_int_logs);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* hipstr_last_error(const hipstr_ctx_t* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

hipstr_status_t hipstr_set_stream(hipstr_ctx_t* ctx, void* cuda_stream) {
  if (!ctx) return HIPSTR_ERR_BAD_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return HIPSTR_OK;
}

int64_t hipstr_batch_num_alignments(const hipstr_align_batch_t* batch) { return batch ? count_alignments(batch) : 0; }
int32_t hipstr_last_launch_count(const hipstr_ctx_t* ctx) { return ctx ? ctx->last_launches : 0; }
void hipstr_trace_seconds(const hipstr_ctx_t* ctx, double* seconds4) {
  if (ctx && seconds4) for (int i = 0; i < 4; i++) seconds4[i] = ctx->trace_seconds[i];
}

void hipstr_last_traffic(const hipstr_ctx_t* ctx, int64_t* h2d, int64_t* d2h, int32_t* launches) {
  if (h2d) *h2d = ctx ? ctx->h2d_bytes : 0;
  if (d2h) *d2h = ctx ? ctx->d2h_bytes : 0;
  if (launches) *launches = ctx ? ctx->last_launches : 0;
}
hipstr_status_t hipstr_enable_timing(hipstr_ctx_t* ctx, int enable) {
  if (!ctx) return HIPSTR_ERR_BAD_ARG;
  ctx->timing = enable != 0;
  return HIPSTR_OK;
}
float hipstr_last_kernel_ms(const hipstr_ctx_t* ctx) { return ctx ? ctx->last_ms : 0.f; }

/* Test/bench hook (not in the public header): waits for the stream, then sums the stage times of
 * every call since the last collect: k1_ms = alignment kernels, rest_ms = scatter + posteriors. */
hipstr_status_t hipstr_collect_timing(hipstr_ctx_t* ctx, double* k1_ms, double* rest_ms, int32_t* n_calls) {
  if (!ctx) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(wait_stream(ctx, ctx->stream));
  double a = 0, b = 0;
  for (auto& ev : ctx->pending) {
    float t0 = 0, t1 = 0;
    CU(cudaEventElapsedTime(&t0, ev.e[0], ev.e[1]));
    CU(cudaEventElapsedTime(&t1, ev.e[1], ev.e[2]));
    a += t0; b += t1;
    ctx->free_events.push_back(ev);
  }
  if (k1_ms) *k1_ms = a;
  if (rest_ms) *rest_ms = b;
  if (n_calls) *n_calls = (int32_t)ctx->pending.size();
  if (!ctx->pending.empty()) ctx->last_ms = (float)(a / ctx->pending.size());
  ctx->pending.clear();
  return HIPSTR_OK;
}

hipstr_status_t hipstr_upload_batch(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch, hipstr_dev_batch_t** out) {
  if (!ctx || !batch || !out) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  hipstr_dev_batch* d = new hipstr_dev_batch();
  hipstr_status_t st = stage(ctx, batch, *d);
  end_call(ctx);
  if (st != HIPSTR_OK) { d->release(); delete d; return st; }
  *out = d;
  return HIPSTR_OK;
}

void hipstr_free_batch(hipstr_ctx_t* ctx, hipstr_dev_batch_t* h) {
  if (!h) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  h->release();
  delete h;
}

hipstr_status_t hipstr_align_batch_dev(hipstr_ctx_t* ctx, const hipstr_dev_batch_t* h, double* ll_dev, int32_t* pos_dev) {
  if (!ctx || !h || !ll_dev) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  StageEvents ev;
  hipstr_status_t st;
  if ((st = mark(ctx, ev, 0)) != HIPSTR_OK) return st;
  if ((st = run_align(ctx, *h, ll_dev, pos_dev)) != HIPSTR_OK) return st;
  if ((st = mark(ctx, ev, 1)) != HIPSTR_OK) return st;
  return mark(ctx, ev, 2);
}

hipstr_status_t hipstr_align_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch, double* ll_out, int32_t* seed_hap_pos) {
  if (!ctx || !batch || !ll_out) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  hipstr_status_t st = stage(ctx, batch, ctx->scratch);
  if (st != HIPSTR_OK) return st;
  const size_t n = (size_t)ctx->scratch.n_out;
  if (n == 0) return HIPSTR_OK;
  CU(ctx->d_ll.reserve(n * sizeof(double)));
  if (seed_hap_pos) CU(ctx->d_pos.reserve(n * sizeof(int32_t)));
  const bool masked = batch->realign_pool || batch->realign_hap;
  if (masked) {   // entries the masks exclude must come back untouched (HapAligner.cpp:326-329,615-619)
    CU(put(ctx->d_ll, ll_out, n, ctx->stream));
    if (seed_hap_pos) CU(put(ctx->d_pos, seed_hap_pos, n, ctx->stream));
  }
  st = run_align(ctx, ctx->scratch, (double*)ctx->d_ll.p, seed_hap_pos ? (int32_t*)ctx->d_pos.p : nullptr);
  if (st != HIPSTR_OK) return st;
  CU(get(ctx, ll_out, ctx->d_ll.p, n));
  if (seed_hap_pos) CU(get(ctx, seed_hap_pos, ctx->d_pos.p, n));
  CU(wait_stream(ctx, ctx->stream));
  end_call(ctx);
  return HIPSTR_OK;
}

/* Test hook (not part of the public header): after the next align call, returns the last-column
 * match values [2][l_max] of job 0's last haplotype, to localise a parity failure. */
hipstr_status_t hipstr_debug_lastcols(hipstr_ctx_t* ctx, int enable, double* out, int32_t n) {
  if (!ctx) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  if (enable && !ctx->d_debug) { CU(cudaMalloc(&ctx->d_debug, 8192 * sizeof(double))); CU(cudaMemset(ctx->d_debug, 0, 8192 * sizeof(double))); }
  if (out && ctx->d_debug) CU(cudaMemcpy(out, ctx->d_debug, (size_t)std::min(n, 8192) * sizeof(double), cudaMemcpyDeviceToHost));
  if (!enable && ctx->d_debug) { cudaFree(ctx->d_debug); ctx->d_debug = nullptr; }
  return HIPSTR_OK;
}

hipstr_status_t hipstr_upload_genotype_batch(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                             const hipstr_reads_batch_t* reads, hipstr_dev_genotype_t** out) {
  if (!ctx || !batch || !reads || !out) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  hipstr_dev_genotype* g = new hipstr_dev_genotype();
  hipstr_status_t st = stage(ctx, batch, g->align);
  if (st == HIPSTR_OK) st = stage_reads(ctx, batch, reads, *g);
  end_call(ctx);
  if (st != HIPSTR_OK) { g->release(); delete g; return st; }
  *out = g;
  return HIPSTR_OK;
}

void hipstr_free_genotype_batch(hipstr_ctx_t* ctx, hipstr_dev_genotype_t* h) {
  if (!h) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  h->release();
  delete h;
}

hipstr_status_t hipstr_genotype_batch_dev(hipstr_ctx_t* ctx, const hipstr_dev_genotype_t* h, const hipstr_genotype_out_t* o) {
  if (!ctx || !h || !o || !o->read_ll || !o->post || !o->sample_ll) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  return run_genotype(ctx, *const_cast<hipstr_dev_genotype*>(h), *o);
}

hipstr_status_t hipstr_genotype_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch,
                                           const hipstr_reads_batch_t* reads, const hipstr_genotype_out_t* o) {
  if (!ctx || !batch || !reads || !o || !o->read_ll || !o->post || !o->sample_ll) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  hipstr_dev_genotype& g = ctx->gscratch;
  hipstr_status_t st = stage(ctx, batch, g.align);
  if (st == HIPSTR_OK) st = stage_reads(ctx, batch, reads, g);
  if (st != HIPSTR_OK) return st;
  DevBuf* d = ctx->d_out;
  CU(d[0].reserve(std::max<size_t>((size_t)g.n_elems * sizeof(double), 16)));
  CU(d[1].reserve(std::max<size_t>((size_t)g.n_reads * sizeof(int32_t), 16)));
  CU(d[2].reserve(std::max<size_t>((size_t)g.post_size * sizeof(double), 16)));
  CU(d[3].reserve(std::max<size_t>((size_t)g.n_samples * sizeof(double), 16)));
  CU(d[4].reserve(std::max<size_t>((size_t)g.n_samples * 2 * sizeof(int32_t), 16)));
  CU(d[5].reserve(std::max<size_t>((size_t)g.n_loci * sizeof(double), 16)));
  if (g.masked) {   // in-place semantics of log_aln_probs_ / seed_positions_ under the masks
    CU(put(d[0], o->read_ll, (size_t)g.n_elems, ctx->stream));
    if (o->read_seed) CU(put(d[1], o->read_seed, (size_t)g.n_reads, ctx->stream));
    // K1 writes only the realigned entries of the pool buffer; K2 reads only those
  }
  hipstr_genotype_out_t dev;
  dev.read_ll = (double*)d[0].p;
  dev.read_seed = o->read_seed ? (int32_t*)d[1].p : nullptr;
  dev.post = (double*)d[2].p;
  dev.sample_ll = (double*)d[3].p;
  dev.best = o->best ? (int32_t*)d[4].p : nullptr;
  dev.total_ll = o->total_ll ? (double*)d[5].p : nullptr;
  st = run_genotype(ctx, g, dev);
  if (st != HIPSTR_OK) return st;
  CU(get(ctx, o->read_ll, d[0].p, (size_t)g.n_elems));
  if (o->read_seed) CU(get(ctx, o->read_seed, d[1].p, (size_t)g.n_reads));
  CU(get(ctx, o->post, d[2].p, (size_t)g.post_size));
  CU(get(ctx, o->sample_ll, d[3].p, (size_t)g.n_samples));
  if (o->best) CU(get(ctx, o->best, d[4].p, (size_t)g.n_samples * 2));
  if (o->total_ll) CU(get(ctx, o->total_ll, d[5].p, (size_t)g.n_loci));
  CU(wait_stream(ctx, ctx->stream));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_scatter_pool_lls_host(hipstr_ctx_t* ctx, int32_t n_reads, int32_t n_haps, const double* pool_ll,
                                             const int32_t* pool_seed, const int32_t* pool_index,
                                             const uint8_t* second_mate, const uint8_t* copy_read,
                                             const uint8_t* realign_hap, double* read_ll, int32_t* read_seed) {
  if (!ctx || n_reads < 0 || n_haps <= 0) return HIPSTR_ERR_BAD_ARG;
  if (n_reads == 0) return HIPSTR_OK;
  if (!pool_ll || !pool_index || !second_mate || !read_ll || (read_seed && !pool_seed)) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  int32_t n_pools = 0;
  for (int r = 0; r < n_reads; r++) {
    if (pool_index[r] < 0) return fail(ctx, HIPSTR_ERR_BAD_ARG, "negative pool index");
    n_pools = std::max(n_pools, pool_index[r] + 1);
  }
  cudaStream_t s = ctx->stream;
  DevBuf* m = ctx->d_misc;
  CU(put(m[0], pool_ll, (size_t)n_pools * n_haps, s));
  CU(put(m[1], pool_index, (size_t)n_reads, s));
  CU(put(m[2], second_mate, (size_t)n_reads, s));
  CU(put(m[3], read_ll, (size_t)n_reads * n_haps, s));
  if (copy_read) CU(put(m[4], copy_read, (size_t)n_reads, s));
  if (realign_hap) CU(put(m[5], realign_hap, (size_t)n_haps, s));
  if (read_seed) { CU(put(m[6], pool_seed, (size_t)n_pools, s)); CU(put(m[7], read_seed, (size_t)n_reads, s)); }
  ScatterParams p;
  p.n_reads = n_reads; p.n_haps = n_haps;
  p.pool_ll = (const double*)m[0].p; p.pool_index = (const int32_t*)m[1].p; p.second_mate = (const uint8_t*)m[2].p;
  p.read_ll = (double*)m[3].p;
  p.copy_read = copy_read ? (const uint8_t*)m[4].p : nullptr;
  p.realign_hap = realign_hap ? (const uint8_t*)m[5].p : nullptr;
  p.pool_seed = read_seed ? (const int32_t*)m[6].p : nullptr;
  p.read_seed = read_seed ? (int32_t*)m[7].p : nullptr;
  CU(launch_scatter(p, s));
  ctx->last_launches = 1;
  CU(get(ctx, read_ll, m[3].p, (size_t)n_reads * n_haps));
  if (read_seed) CU(get(ctx, read_seed, m[7].p, (size_t)n_reads));
  CU(wait_stream(ctx, s));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_posteriors_host(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* locus_read_off,
                                       const int32_t* locus_sample_off, const int32_t* n_haps, const uint8_t* haploid,
                                       const double* read_ll, const double* log_p1, const double* log_p2,
                                       const int32_t* sample_label, const int32_t* read_weight, double* post_out,
                                       double* sample_ll_out, int32_t* best_out, double* total_ll_out) {
  if (!ctx || n_loci < 0) return HIPSTR_ERR_BAD_ARG;
  if (n_loci == 0) return HIPSTR_OK;
  if (!locus_read_off || !locus_sample_off || !n_haps || !haploid || !read_ll || !log_p1 || !log_p2 || !sample_label ||
      !read_weight || !post_out || !sample_ll_out)
    return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  const int32_t R = locus_read_off[n_loci], S = locus_sample_off[n_loci];
  std::vector<PostSample> samples((size_t)S);
  int64_t ll_off = 0, post_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l];
    if (H <= 0 || H + 1 >= 10000) return fail(ctx, HIPSTR_ERR_BAD_ARG, "bad haplotype count");
    const int s0 = locus_sample_off[l], s1 = locus_sample_off[l + 1];
    int r = locus_read_off[l];
    const int r_end = locus_read_off[l + 1];
    for (int s = s0; s < s1; s++) {
      PostSample& ps = samples[s];
      ps.read0 = r;
      while (r < r_end && sample_label[r] == s - s0) r++;   // reads are sample-major (genotyper.h:104-112)
      ps.read1 = r;
      ps.n_haps = H;
      ps.haploid = haploid[l];
      ps.ll_off = ll_off;
      ps.locus_read0 = locus_read_off[l];
      ps.locus = l;
      ps.post_off = post_off + (int64_t)(s - s0) * H * H;
    }
    if (r != r_end) return fail(ctx, HIPSTR_ERR_BAD_ARG, "reads are not sample-major or a label is out of range");
    ll_off += (int64_t)(r_end - locus_read_off[l]) * H;
    post_off += (int64_t)(s1 - s0) * H * H;
  }
  cudaStream_t s = ctx->stream;
  DevBuf* m = ctx->d_misc;
  CU(put(m[0], samples, s));
  CU(put(m[1], locus_sample_off, (size_t)n_loci + 1, s));
  CU(put(m[2], read_ll, (size_t)ll_off, s));
  CU(put(m[3], log_p1, (size_t)R, s));
  CU(put(m[4], log_p2, (size_t)R, s));
  CU(put(m[5], read_weight, (size_t)R, s));
  CU(m[6].reserve(std::max<size_t>((size_t)post_off * sizeof(double), 16)));
  CU(m[7].reserve(std::max<size_t>((size_t)S * sizeof(double), 16)));
  CU(m[8].reserve(std::max<size_t>((size_t)S * 2 * sizeof(int32_t), 16)));
  CU(m[9].reserve(std::max<size_t>((size_t)n_loci * sizeof(double), 16)));
  PostParams p;
  p.n_samples = S; p.n_loci = n_loci;
  p.samples = (const PostSample*)m[0].p; p.locus_sample_off = (const int32_t*)m[1].p;
  p.read_ll = (const double*)m[2].p; p.log_p1 = (const double*)m[3].p; p.log_p2 = (const double*)m[4].p;
  p.read_weight = (const int32_t*)m[5].p;
  p.int_logs = ctx->d_int_logs; p.log_one_half = host_tables().log_one_half;
  p.post_out = (double*)m[6].p; p.sample_ll_out = (double*)m[7].p;
  p.best_out = best_out ? (int32_t*)m[8].p : nullptr;
  p.total_ll_out = total_ll_out ? (double*)m[9].p : nullptr;
  CU(launch_posteriors(p, s));
  ctx->last_launches = total_ll_out ? 2 : 1;
  CU(get(ctx, post_out, m[6].p, (size_t)post_off));
  CU(get(ctx, sample_ll_out, m[7].p, (size_t)S));
  if (best_out) CU(get(ctx, best_out, m[8].p, (size_t)S * 2));
  if (total_ll_out) CU(get(ctx, total_ll_out, m[9].p, (size_t)n_loci));
  CU(wait_stream(ctx, s));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_em_train_host(hipstr_ctx_t* ctx, const hipstr_em_batch_t* b, int32_t max_iter, double min_abs,
                                     double min_frac, double* params_out, uint8_t* converged_out, int32_t* iters_out,
                                     double* ll_out) {
  if (!ctx || !b || b->n_loci < 0 || !params_out || !converged_out) return HIPSTR_ERR_BAD_ARG;
  if (b->n_loci == 0) return HIPSTR_OK;
  if (!b->locus_read_off || !b->locus_sample_off || !b->num_bps || !b->sample_label || !b->log_p1 || !b->log_p2 ||
      !b->motif_len || !b->ref_allele || !b->haploid)
    return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  const int n_loci = b->n_loci;
  const int32_t R = b->locus_read_off[n_loci], S = b->locus_sample_off[n_loci];
  std::vector<EmLocus> loci((size_t)n_loci);
  std::vector<int32_t> allele_of((size_t)R), sample_read_off((size_t)S + 1), bps;
  std::vector<double> gt_prior, params((size_t)n_loci * 6);
  int64_t post_off = 0, row_off = 0;
  int max_alleles = 1;
  for (int l = 0; l < n_loci; l++) {
    const int r0 = b->locus_read_off[l], r1 = b->locus_read_off[l + 1];
    const int s0 = b->locus_sample_off[l], Sl = b->locus_sample_off[l + 1] - s0;
    if (b->motif_len[l] < 1 || b->motif_len[l] > 9) return fail(ctx, HIPSTR_ERR_BAD_ARG, "motif length must be 1..9");
    // allele sizes: reference first, the rest ascending (em_stutter_genotyper.h:59-81)
    std::vector<int32_t> sizes(b->num_bps + r0, b->num_bps + r1);
    std::sort(sizes.begin(), sizes.end());
    sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
    sizes.erase(std::remove(sizes.begin(), sizes.end(), b->ref_allele[l]), sizes.end());
    sizes.insert(sizes.begin(), b->ref_allele[l]);
    const int A = (int)sizes.size();
    if (A > 100) return fail(ctx, HIPSTR_ERR_UNSUPPORTED, "more than 100 distinct STR sizes at one locus");
    max_alleles = std::max(max_alleles, A);
    EmLocus& L = loci[l];
    L.read0 = r0; L.n_reads = r1 - r0; L.sample0 = s0; L.n_samples = Sl;
    L.n_alleles = A; L.period = b->motif_len[l]; L.haploid = b->haploid[l];
    L.allele_off = (int32_t)bps.size();
    L.post_off = post_off; L.row_off = row_off;
    post_off += (int64_t)Sl * A * A;
    row_off += (int64_t)Sl * A;
    // reads: allele index, per-sample ranges; init_log_gt_priors (em_stutter_genotyper.cpp:10-19) with the host libm
    std::vector<int> per_sample((size_t)Sl, 0);
    int at = r0;
    for (int s = 0; s < Sl; s++) {
      sample_read_off[s0 + s] = at;
      while (at < r1 && b->sample_label[at] == s) at++;
      per_sample[s] = at - sample_read_off[s0 + s];
    }
    if (at != r1) return fail(ctx, HIPSTR_ERR_BAD_ARG, "reads are not sample-major or a sample label is out of range");
    std::vector<double> prior((size_t)A, 1.0);
    for (int r = r0; r < r1; r++) {
      const int a = (int)(std::lower_bound(sizes.begin() + 1, sizes.end(), b->num_bps[r]) - sizes.begin());
      allele_of[r] = (b->num_bps[r] == b->ref_allele[l]) ? 0 : a;
      prior[allele_of[r]] += 1.0 / per_sample[b->sample_label[r]];
    }
    double total = 0.0;
    for (int a = 0; a < A; a++) total += prior[a];
    const double log_total = std::log(total);
    for (int a = 0; a < A; a++) gt_prior.push_back(std::log(prior[a]) - log_total);
    bps.insert(bps.end(), sizes.begin(), sizes.end());
    const double init[6] = {0.9, 0.1, 0.1, 0.8, 0.01, 0.01};   // init_stutter_model (:58-61)
    std::copy(init, init + 6, params.begin() + 6 * (size_t)l);
  }
  sample_read_off[S] = R;
  cudaStream_t st = ctx->stream;
  DevBuf* m = ctx->d_misc;
  DevBuf* o = ctx->d_out;
  CU(put(m[0], loci, st));
  CU(put(m[1], allele_of, st));
  CU(put(m[2], b->sample_label, (size_t)R, st));
  CU(put(m[3], sample_read_off, st));
  CU(put(m[4], b->log_p1, (size_t)R, st));
  CU(put(m[5], b->log_p2, (size_t)R, st));
  CU(put(m[6], bps, st));
  CU(put(m[7], gt_prior, st));
  CU(put(m[8], params, st));
  CU(o[0].reserve(std::max<size_t>((size_t)post_off * sizeof(double), 16)));
  CU(o[1].reserve(std::max<size_t>((size_t)row_off * sizeof(double), 16)));
  CU(o[2].reserve(std::max<size_t>((size_t)S * sizeof(double), 16)));
  CU(o[3].reserve((size_t)n_loci));
  CU(o[4].reserve((size_t)n_loci * sizeof(int32_t)));
  CU(o[5].reserve((size_t)n_loci * sizeof(double)));
  EmParams p;
  p.n_loci = n_loci; p.max_iter = max_iter; p.min_abs = min_abs; p.min_frac = min_frac;
  p.loci = (const EmLocus*)m[0].p; p.allele_of = (const int32_t*)m[1].p; p.sample_label = (const int32_t*)m[2].p;
  p.sample_read_off = (const int32_t*)m[3].p; p.log_p1 = (const double*)m[4].p; p.log_p2 = (const double*)m[5].p;
  p.bps = (const int32_t*)m[6].p; p.gt_prior = (double*)m[7].p; p.params = (double*)m[8].p;
  p.int_logs = ctx->d_int_logs; p.log_one_half = host_tables().log_one_half;
  p.post = (double*)o[0].p; p.rowlse = (double*)o[1].p; p.sample_ll = (double*)o[2].p;
  p.converged = (uint8_t*)o[3].p; p.iters = (int32_t*)o[4].p; p.final_ll = (double*)o[5].p;
  CU(launch_em(p, max_alleles, st));
  ctx->last_launches = 1;
  CU(get(ctx, params_out, m[8].p, (size_t)n_loci * 6));
  CU(get(ctx, converged_out, o[3].p, (size_t)n_loci));
  if (iters_out) CU(get(ctx, iters_out, o[4].p, (size_t)n_loci));
  if (ll_out) CU(get(ctx, ll_out, o[5].p, (size_t)n_loci));
  CU(wait_stream(ctx, st));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_extract_genotypes_host(hipstr_ctx_t* ctx, int32_t n_loci, const int32_t* locus_sample_off,
                                              const int32_t* n_haps, const int32_t* n_variants,
                                              const int32_t* hap_to_allele, const uint8_t* haploid, const double* post,
                                              const double* sample_ll, int32_t* best_hap, int32_t* best_gt,
                                              double* log_phased, double* log_unphased, double* hap_log_phased,
                                              double* hap_log_unphased, double* gl, double* phased_gl, double* gl_diff,
                                              int32_t* pl) {
  if (!ctx || n_loci < 0) return HIPSTR_ERR_BAD_ARG;
  if (n_loci == 0) return HIPSTR_OK;
  if (!locus_sample_off || !n_haps || !n_variants || !hap_to_allele || !haploid || !post || !sample_ll || !best_hap ||
      !best_gt || !log_phased || !log_unphased || !hap_log_phased || !hap_log_unphased || !gl || !phased_gl || !gl_diff || !pl)
    return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  const int32_t S = locus_sample_off[n_loci];
  std::vector<ExtractSample> samples((size_t)S);
  int64_t post_off = 0, gl_off = 0, pgl_off = 0;
  int32_t h2a_off = 0;
  for (int l = 0; l < n_loci; l++) {
    const int H = n_haps[l], V = n_variants[l];
    if (H <= 0 || V <= 0 || V > H || H + 1 >= 10000) return fail(ctx, HIPSTR_ERR_BAD_ARG, "bad haplotype / allele count");
    for (int h = 0; h < H; h++)
      if (hap_to_allele[h2a_off + h] < 0 || hap_to_allele[h2a_off + h] >= V) return fail(ctx, HIPSTR_ERR_BAD_ARG, "hap_to_allele out of range");
    const int G = haploid[l] ? V : V * (V + 1) / 2, PG = haploid[l] ? V : V * V;
    for (int s = locus_sample_off[l]; s < locus_sample_off[l + 1]; s++) {
      ExtractSample& e = samples[s];
      e.n_haps = H; e.n_variants = V; e.haploid = haploid[l]; e.h2a_off = h2a_off;
      e.post_off = post_off; e.gl_off = gl_off; e.pgl_off = pgl_off;
      post_off += (int64_t)H * H; gl_off += G; pgl_off += PG;
    }
    h2a_off += H;
  }
  cudaStream_t st = ctx->stream;
  DevBuf* m = ctx->d_misc;
  DevBuf* o = ctx->d_out;
  CU(put(m[0], samples, st));
  CU(put(m[1], hap_to_allele, (size_t)h2a_off, st));
  CU(put(m[2], post, (size_t)post_off, st));
  CU(put(m[3], sample_ll, (size_t)S, st));
  CU(m[4].reserve(std::max<size_t>((size_t)S * 4 * sizeof(int32_t), 16)));
  CU(m[5].reserve(std::max<size_t>((size_t)S * 5 * sizeof(double), 16)));
  CU(o[0].reserve(std::max<size_t>((size_t)gl_off * sizeof(double), 16)));
  CU(o[1].reserve(std::max<size_t>((size_t)pgl_off * sizeof(double), 16)));
  CU(o[2].reserve(std::max<size_t>((size_t)gl_off * sizeof(int32_t), 16)));
  ExtractParams p;
  p.n_samples = S; p.samples = (const ExtractSample*)m[0].p; p.hap_to_allele = (const int32_t*)m[1].p;
  p.post = (const double*)m[2].p; p.sample_ll = (const double*)m[3].p; p.int_logs = ctx->d_int_logs;
  int32_t* di = (int32_t*)m[4].p;
  double* dd = (double*)m[5].p;
  p.best_hap = di; p.best_gt = di + 2 * (size_t)S;
  p.log_phased = dd; p.log_unphased = dd + S; p.hap_log_phased = dd + 2 * (size_t)S; p.hap_log_unphased = dd + 3 * (size_t)S;
  p.gl_diff = dd + 4 * (size_t)S;
  p.gl = (double*)o[0].p; p.phased_gl = (double*)o[1].p; p.pl = (int32_t*)o[2].p;
  CU(launch_extract(p, st));
  ctx->last_launches = 1;
  CU(get(ctx, best_hap, p.best_hap, (size_t)S * 2));
  CU(get(ctx, best_gt, p.best_gt, (size_t)S * 2));
  CU(get(ctx, log_phased, p.log_phased, (size_t)S));
  CU(get(ctx, log_unphased, p.log_unphased, (size_t)S));
  CU(get(ctx, hap_log_phased, p.hap_log_phased, (size_t)S));
  CU(get(ctx, hap_log_unphased, p.hap_log_unphased, (size_t)S));
  CU(get(ctx, gl_diff, p.gl_diff, (size_t)S));
  CU(get(ctx, gl, p.gl, (size_t)gl_off));
  CU(get(ctx, phased_gl, p.phased_gl, (size_t)pgl_off));
  CU(get(ctx, pl, p.pl, (size_t)gl_off));
  CU(wait_stream(ctx, st));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_nw_align_batch_host(hipstr_ctx_t* ctx, int32_t n_pairs, const int32_t* ref_off, const char* ref_seqs,
                                           const int32_t* read_off, const char* read_seqs, int32_t use_ref_end_penalty,
                                           int32_t ops_stride, char* ops, int32_t* ops_len, float* score) {
  if (!ctx || n_pairs < 0) return HIPSTR_ERR_BAD_ARG;
  if (n_pairs == 0) return HIPSTR_OK;
  if (!ref_off || !ref_seqs || !read_off || !read_seqs || !ops || !ops_len || !score) return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  int max_ref = 1, max_read = 1;
  for (int i = 0; i < n_pairs; i++) {
    const int l1 = ref_off[i + 1] - ref_off[i], l2 = read_off[i + 1] - read_off[i];
    if (l1 < 1 || l2 < 1) return fail(ctx, HIPSTR_ERR_BAD_ARG, "empty sequence in an alignment pair");
    max_ref = std::max(max_ref, l1);
    max_read = std::max(max_read, l2);
  }
  if (ops_stride < max_ref + max_read + 1) return fail(ctx, HIPSTR_ERR_BAD_ARG, "ops_stride too small");
  if (nw_shared_bytes(max_ref, max_read) > (size_t)200 * 1024)
    return fail(ctx, HIPSTR_ERR_UNSUPPORTED, "window too long for the shared-memory rows");
  cudaStream_t s = ctx->stream;
  DevBuf* m = ctx->d_misc;
  DevBuf* o = ctx->d_out;
  // one trace byte per cell, in global memory; pairs are processed in chunks of at most 2 GB of trace
  std::vector<int64_t> trace_off((size_t)n_pairs + 1, 0);
  for (int i = 0; i < n_pairs; i++)
    trace_off[i + 1] = trace_off[i] + (int64_t)((ref_off[i + 1] - ref_off[i] + 15) / 16 * 16) * (read_off[i + 1] - read_off[i]);
  const int64_t kTraceBudget = (int64_t)2 << 30;
  int64_t widest = 0;
  for (int first = 0; first < n_pairs;) {
    int last = first;
    while (last < n_pairs && trace_off[last + 1] - trace_off[first] <= kTraceBudget) last++;
    if (last == first) return fail(ctx, HIPSTR_ERR_UNSUPPORTED, "one alignment exceeds the trace budget");
    widest = std::max(widest, trace_off[last] - trace_off[first]);
    first = last;
  }
  CU(put(m[0], ref_off, (size_t)n_pairs + 1, s));
  CU(put(m[1], ref_seqs, (size_t)ref_off[n_pairs], s));
  CU(put(m[2], read_off, (size_t)n_pairs + 1, s));
  CU(put(m[3], read_seqs, (size_t)read_off[n_pairs], s));
  CU(put(m[4], trace_off, s));
  const size_t T = (size_t)n_pairs;
  CU(o[0].reserve(T * ops_stride));
  CU(o[1].reserve(T * sizeof(int32_t)));
  CU(o[2].reserve(T * sizeof(float)));
  CU(o[3].reserve(T * 2 * sizeof(int32_t)));
  CU(ctx->d_last.reserve((size_t)std::max<int64_t>(widest, 16)));
  NwParams p;
  std::memset(&p, 0, sizeof(p));
  p.ref_off = (const int32_t*)m[0].p; p.ref_seqs = (const char*)m[1].p;
  p.read_off = (const int32_t*)m[2].p; p.read_seqs = (const char*)m[3].p;
  p.use_ref_end_penalty = use_ref_end_penalty ? 1 : 0;
  p.max_ref = max_ref; p.max_read = max_read; p.ops_stride = ops_stride;
  p.out_ops = (char*)o[0].p; p.out_len = (int32_t*)o[1].p; p.out_score = (float*)o[2].p;
  p.trace_off = (const int64_t*)m[4].p; p.trace = (unsigned char*)ctx->d_last.p; p.end_cell = (int32_t*)o[3].p;
  int n_sm = 148;
  CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
  int launches = 0;
  for (int first = 0; first < n_pairs;) {
    int last = first;
    while (last < n_pairs && trace_off[last + 1] - trace_off[first] <= kTraceBudget) last++;
    p.first_pair = first;
    p.n_pairs = last - first;
    // one warp per CTA; up to 32 of them per SM, and a few waves keep the tail short
    CU(launch_nw(p, n_sm * 64, s));
    launches += 2;
    first = last;
  }
  ctx->last_launches = launches;
  CU(get(ctx, ops, p.out_ops, T * ops_stride));
  CU(get(ctx, ops_len, p.out_len, T));
  CU(get(ctx, score, p.out_score, T));
  CU(wait_stream(ctx, s));
  end_call(ctx);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_snp_phasing_batch_host(hipstr_ctx_t* ctx, const hipstr_snp_phasing_t* b, double* log_p1, double* log_p2,
                                              int32_t* counts) {
  if (!ctx || !b || b->n_entries < 0 || b->n_alns < 0 || b->n_sets < 0) return HIPSTR_ERR_BAD_ARG;
  if (b->n_entries == 0) return HIPSTR_OK;
  if (!log_p1 || !log_p2 || !counts || !b->entry_aln_off || !b->entry_snp_set) return HIPSTR_ERR_BAD_ARG;
  if (b->n_alns > 0 && (!b->aln_pos || !b->aln_end || !b->aln_seq_off || !b->bases || !b->quals || !b->aln_cigar_off ||
                        !b->cigar_type || !b->cigar_len))
    return HIPSTR_ERR_BAD_ARG;
  if (b->n_sets > 0 && (!b->set_off || (b->set_off[b->n_sets] > 0 && (!b->snp_pos || !b->snp_base1 || !b->snp_base2))))
    return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  if (b->entry_aln_off[0] != 0 || b->entry_aln_off[b->n_entries] != b->n_alns)
    return fail(ctx, HIPSTR_ERR_BAD_ARG, "entry_aln_off does not cover the alignments");
  for (int e = 0; e < b->n_entries; e++) {
    if (b->entry_aln_off[e + 1] < b->entry_aln_off[e]) return fail(ctx, HIPSTR_ERR_BAD_ARG, "entry_aln_off not ascending");
    if (b->entry_snp_set[e] < -1 || b->entry_snp_set[e] >= b->n_sets) return fail(ctx, HIPSTR_ERR_BAD_ARG, "SNP set out of range");
  }
  for (int s = 0; s < b->n_sets; s++)
    for (int i = b->set_off[s] + 1; i < b->set_off[s + 1]; i++)
      if (b->snp_pos[i] < b->snp_pos[i - 1]) return fail(ctx, HIPSTR_ERR_BAD_ARG, "SNP positions of a set must be non-decreasing");
  cudaStream_t s = ctx->stream;
  DevBuf* m = ctx->d_misc;
  DevBuf* o = ctx->d_out;
  const size_t E = (size_t)b->n_entries, A = (size_t)b->n_alns;
  const size_t n_snps = b->n_sets ? (size_t)b->set_off[b->n_sets] : 0;
  const int32_t zero = 0;
  CU(put(m[0], b->entry_aln_off, E + 1, s));
  CU(put(m[1], b->entry_snp_set, E, s));
  CU(put(m[2], b->aln_pos ? b->aln_pos : &zero, std::max<size_t>(A, 1), s));
  CU(put(m[3], b->aln_end ? b->aln_end : &zero, std::max<size_t>(A, 1), s));
  CU(put(m[4], b->aln_seq_off ? b->aln_seq_off : &zero, A ? A + 1 : 1, s));
  const size_t n_bases = A ? (size_t)b->aln_seq_off[A] : 0, n_ops = A ? (size_t)b->aln_cigar_off[A] : 0;
  const char pad = 0;
  CU(put(m[5], n_bases ? b->bases : &pad, std::max<size_t>(n_bases, 1), s));
  CU(put(m[6], n_bases ? b->quals : &pad, std::max<size_t>(n_bases, 1), s));
  CU(put(m[7], b->aln_cigar_off ? b->aln_cigar_off : &zero, A ? A + 1 : 1, s));
  CU(put(m[8], n_ops ? b->cigar_type : &pad, std::max<size_t>(n_ops, 1), s));
  CU(put(m[9], n_ops ? b->cigar_len : &zero, std::max<size_t>(n_ops, 1), s));
  CU(put(m[10], b->n_sets ? b->set_off : &zero, (size_t)b->n_sets + 1, s));
  // the three SNP arrays share one buffer: positions, then the two allele bytes
  std::vector<unsigned char> snps(std::max<size_t>(n_snps, 1) * 6);
  if (n_snps) {
    std::memcpy(snps.data(), b->snp_pos, n_snps * 4);
    std::memcpy(snps.data() + n_snps * 4, b->snp_base1, n_snps);
    std::memcpy(snps.data() + n_snps * 5, b->snp_base2, n_snps);
  }
  CU(put(m[11], snps.data(), snps.size(), s));
  CU(o[0].reserve(E * sizeof(double)));
  CU(o[1].reserve(E * sizeof(double)));
  CU(o[2].reserve(E * 4 * sizeof(int32_t)));
  SnpPhaseParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_entries = b->n_entries;
  p.entry_aln_off = (const int32_t*)m[0].p; p.entry_snp_set = (const int32_t*)m[1].p;
  p.aln_pos = (const int32_t*)m[2].p; p.aln_end = (const int32_t*)m[3].p; p.aln_seq_off = (const int32_t*)m[4].p;
  p.bases = (const char*)m[5].p; p.quals = (const char*)m[6].p;
  p.aln_cigar_off = (const int32_t*)m[7].p; p.cigar_type = (const char*)m[8].p; p.cigar_len = (const int32_t*)m[9].p;
  p.set_off = (const int32_t*)m[10].p;
  p.snp_pos = (const uint32_t*)m[11].p;
  p.snp_base1 = (const char*)m[11].p + n_snps * 4;
  p.snp_base2 = (const char*)m[11].p + n_snps * 5;
  p.qual_lut = ctx->d_qual_lut;
  p.out_log_p1 = (double*)o[0].p; p.out_log_p2 = (double*)o[1].p; p.out_counts = (int32_t*)o[2].p;
  int n_sm = 148;
  CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
  cudaEvent_t t0 = nullptr, t1 = nullptr;   // kernel time on the launching stream, read by hipstr_last_kernel_ms
  if (ctx->timing) { CU(cudaEventCreate(&t0)); CU(cudaEventCreate(&t1)); CU(cudaEventRecord(t0, s)); }
  CU(launch_snp_phase(p, n_sm, s));
  if (ctx->timing) CU(cudaEventRecord(t1, s));
  ctx->last_launches = 1;
  CU(get(ctx, log_p1, p.out_log_p1, E));
  CU(get(ctx, log_p2, p.out_log_p2, E));
  CU(get(ctx, counts, p.out_counts, E * 4));
  CU(wait_stream(ctx, s));
  if (ctx->timing) {
    CU(cudaEventElapsedTime(&ctx->last_ms, t0, t1));
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
  }
  end_call(ctx);
  for (size_t e = 0; e < E; e++)
    if (counts[4 * e + 3] != 0) return fail(ctx, HIPSTR_ERR_BAD_ARG, "an alignment's CIGAR is invalid or inconsistent with its bases");
  return HIPSTR_OK;
}

hipstr_status_t hipstr_trace_batch_host(hipstr_ctx_t* ctx, const hipstr_align_batch_t* batch, const int32_t* block_start,
                                        int32_t n_traces, const int32_t* trace_pool, const int32_t* trace_hap,
                                        const hipstr_trace_out_t* out) {
  if (!ctx || !batch || !block_start || n_traces < 0 || !out) return HIPSTR_ERR_BAD_ARG;
  if (n_traces == 0) return HIPSTR_OK;
  if (!trace_pool || !trace_hap || !out->hap_aln || !out->seed_hap_pos || !out->stutter_size || !out->span_start ||
      !out->span_len || !out->flank_ins || !out->flank_del || !out->n_indels || !out->indels || !out->n_snps || !out->snps)
    return HIPSTR_ERR_BAD_ARG;
  CU(cudaSetDevice(ctx->device));
  begin_call(ctx);
  // a trace sees its haplotype from scratch: lower with fresh homopolymer classes, no masks
  hipstr_align_batch_t plain = *batch;
  plain.realign_pool = nullptr;
  plain.realign_hap = nullptr;
  FlatBatch& f = ctx->flat;
  std::string err;
  hipstr_status_t st;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_mark = now();
  try {
    st = flatten_batch(&plain, f, err, /*fresh_rows=*/true);
  } catch (const std::bad_alloc&) {
    return fail(ctx, HIPSTR_ERR_CUDA, "out of host memory while staging the batch");
  }
  if (st != HIPSTR_OK) return fail(ctx, st, err);
  ctx->trace_seconds[0] += now() - t_mark; t_mark = now();
  int n_max = 1, l_max = 1;
  std::vector<int32_t> block_ref_end((size_t)batch->n_blocks), locus_block0((size_t)batch->n_loci);
  for (int l = 0; l < batch->n_loci; l++) {
    locus_block0[l] = batch->locus_block_off[l];
    for (int b = batch->locus_block_off[l]; b < batch->locus_block_off[l + 1]; b++) {
      const int o0 = batch->block_opt_off[b];
      block_ref_end[b] = block_start[b] + (batch->opt_seq_off[o0 + 1] - batch->opt_seq_off[o0]);
    }
  }
  for (int t = 0; t < n_traces; t++) {
    const int p = trace_pool[t];
    if (p < 0 || p >= batch->n_pools) return fail(ctx, HIPSTR_ERR_BAD_ARG, "trace_pool out of range");
    const DevPool& dp = f.pools[p];
    if (trace_hap[t] < 0 || trace_hap[t] >= dp.n_haps) return fail(ctx, HIPSTR_ERR_BAD_ARG, "trace_hap out of range");
    if (dp.seed <= 0 || dp.seed >= dp.len - 1) return fail(ctx, HIPSTR_ERR_INVALID_SEED, "a traced read needs a seed");
    n_max = std::max(n_max, dp.len);
    l_max = std::max(l_max, f.hapsides[dp.hap_rec0 + 2 * trace_hap[t]].len);
  }
  if (out->aln_stride < n_max + l_max + 2) return fail(ctx, HIPSTR_ERR_BAD_ARG, "aln_stride too small");
  cudaStream_t s = ctx->stream;
  hipstr_dev_batch& d = ctx->scratch;
  CU(put(d.pools, f.pools, s));
  CU(put(d.bases, f.bases, s));
  CU(put(d.quals, f.quals, s));
  CU(put(d.hapsides, f.hapsides, s));
  CU(put(d.hapbytes, f.hapbytes, s));
  CU(put(d.blocks, f.blocks, s));
  CU(put(d.reps, f.reps, s));
  CU(put(d.progs, f.progs, s));
  CU(put(d.tabs, f.rep_tabs, s));
  CU(put(d.slot_reps, f.slot_reps, s));
  DevBuf* m = ctx->d_misc;
  DevBuf* o = ctx->d_out;
  CU(put(m[0], trace_pool, (size_t)n_traces, s));
  CU(put(m[1], trace_hap, (size_t)n_traces, s));
  CU(put(m[2], block_start, (size_t)batch->n_blocks, s));
  CU(put(m[3], block_ref_end, s));
  CU(put(m[4], locus_block0, s));
  // The forward pass is the alignment kernel pair in TRACE mode: one K1b job per trace = (pool, ONE haplotype), bucketed
  // by columns per lane like the alignment jobs, and one K1a job per (trace, repeat block of its haplotype).  Every
  // trace owns a slab of stutter tables (one per repeat block), of predecessor bytes and of artifact tables.
  std::vector<DevJob> jobs[kNumColVariants];
  std::vector<int64_t> job_t_off[kNumColVariants], stut_t_off, dec_off((size_t)n_traces), art_off((size_t)n_traces);
  std::vector<DevStutJob> stut_jobs;
  int64_t t_doubles = 0, dec_bytes = 0, art_ints = 0;
  int stut_n_max = 16, n_max_v[kNumColVariants], l_all = 2;
  for (int v = 0; v < kNumColVariants; v++) n_max_v[v] = 16;
  {
    const std::vector<int32_t>& locus_slot0 = f.locus_slot0;   // DevBlock::tslot is local to the locus
    for (int t = 0; t < n_traces; t++) {
      const DevPool& dp = f.pools[trace_pool[t]];
      const DevHapSide& hs = f.hapsides[dp.hap_rec0 + 2 * trace_hap[t]];
      int v = -1;
      for (int c = 0; c < kNumColVariants && v < 0; c++) {
        const int C = kColVariants[c];
        if ((dp.seed + C - 1) / C + (dp.len - dp.seed - 1 + C - 1) / C <= 32) v = c;
      }
      if (v < 0) return fail(ctx, HIPSTR_ERR_UNSUPPORTED, "read longer than the kernel's limit");
      const int pitch = hipstr_t_pitch(dp.len);
      int n_rep = 0;
      for (int b = 0; b < hs.n_blocks; b++) {
        const DevBlock& db = f.blocks[hs.blk_off + b];
        if (db.rep < 0) continue;
        DevStutJob sj = {trace_pool[t], locus_slot0[dp.locus] + db.tslot, 1, n_rep};
        stut_jobs.push_back(sj);
        stut_t_off.push_back(t_doubles);
        n_rep++;
      }
      DevJob j = {trace_pool[t], trace_hap[t], trace_hap[t] + 1, t};
      jobs[v].push_back(j);
      job_t_off[v].push_back(t_doubles);
      t_doubles += (int64_t)std::max(n_rep, 1) * HIPSTR_NUM_ARTIFACTS * pitch;
      dec_off[t] = dec_bytes;
      dec_bytes += (int64_t)hs.len * (dp.len - 1);
      art_off[t] = art_ints;
      art_ints += (int64_t)2 * hs.n_blocks * (dp.len - 1);
      stut_n_max = std::max(stut_n_max, (dp.len + 15) / 16 * 16);
      n_max_v[v] = std::max(n_max_v[v], (dp.len + 3) / 4 * 4);
      l_all = std::max(l_all, (hs.len + 1) / 2 * 2);
    }
  }
  CU(put(m[5], stut_jobs, s));
  CU(put(m[6], stut_t_off, s));
  CU(put(m[7], dec_off, s));
  CU(put(m[8], art_off, s));
  CU(ctx->d_stut.reserve((size_t)std::max<int64_t>(t_doubles, 2) * sizeof(double)));
  CU(ctx->d_stut_pos.reserve((size_t)std::max<int64_t>(t_doubles, 2) * sizeof(int32_t)));
  CU(ctx->d_dec.reserve((size_t)std::max<int64_t>(dec_bytes, 16)));
  CU(ctx->d_art.reserve((size_t)std::max<int64_t>(art_ints, 4) * sizeof(int32_t)));
  CU(ctx->d_last.reserve((size_t)HIPSTR_MAX_ALIGN_CTAS * HIPSTR_WARPS_PER_CTA * 2 * l_all * sizeof(double)));
  CU(ctx->d_counters.reserve((kNumColVariants + 1) * sizeof(int32_t)));
  CU(cudaMemsetAsync(ctx->d_counters.p, 0, (kNumColVariants + 1) * sizeof(int32_t), s));
  const size_t T = (size_t)n_traces;
  CU(o[0].reserve(T * out->aln_stride));
  CU(o[1].reserve(T * (1 + 3 * HIPSTR_MAX_BLOCKS_PER_LOCUS + 4 + 2 * HIPSTR_MAX_TRACE_INDELS + 2 * HIPSTR_MAX_TRACE_SNPS) * sizeof(int32_t)));
  CU(o[2].reserve(T * sizeof(int32_t)));
  int32_t* di = (int32_t*)o[1].p;
  TraceWalkParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_traces = n_traces;
  p.trace_pool = (const int32_t*)m[0].p; p.trace_hap = (const int32_t*)m[1].p;
  p.pools = (const DevPool*)d.pools.p; p.bases = (const char*)d.bases.p; p.quals = (const char*)d.quals.p;
  p.hapsides = (const DevHapSide*)d.hapsides.p; p.hapbytes = (const uint8_t*)d.hapbytes.p;
  p.blocks = (const DevBlock*)d.blocks.p;
  p.qual_lut = ctx->d_qual_lut;
  p.block_start = (const int32_t*)m[2].p; p.block_ref_end = (const int32_t*)m[3].p; p.locus_block0 = (const int32_t*)m[4].p;
  p.dec = (const unsigned char*)ctx->d_dec.p; p.dec_off = (const int64_t*)m[7].p;
  p.art = (const int32_t*)ctx->d_art.p; p.art_off = (const int64_t*)m[8].p;
  p.seed_pos = (const int32_t*)o[2].p;
  p.aln_stride = out->aln_stride;
  p.out_aln = (char*)o[0].p;
  p.out_seed_pos = di; di += T;
  p.out_stutter = di; di += T * HIPSTR_MAX_BLOCKS_PER_LOCUS;
  p.out_span_start = di; di += T * HIPSTR_MAX_BLOCKS_PER_LOCUS;
  p.out_span_len = di; di += T * HIPSTR_MAX_BLOCKS_PER_LOCUS;
  p.out_flank_ins = di; di += T;
  p.out_flank_del = di; di += T;
  p.out_n_indels = di; di += T;
  p.out_n_snps = di; di += T;
  p.out_indels = di; di += T * 2 * HIPSTR_MAX_TRACE_INDELS;
  p.out_snps = di;
  CU(cudaMemsetAsync(o[0].p, 0, T * out->aln_stride, s));
  // K1a in TRACE mode: tables + best artifact positions
  StutParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.jobs = (const DevStutJob*)m[5].p; sp.n_jobs = (int32_t)stut_jobs.size(); sp.n_max = stut_n_max;
  sp.pools = p.pools; sp.bases = p.bases; sp.quals = p.quals;
  sp.slot_reps = (const DevSlotReps*)d.slot_reps.p; sp.reps = (const DevRep*)d.reps.p;
  sp.progs = (const DevProgEntry*)d.progs.p; sp.rep_tabs = (const int32_t*)d.tabs.p;
  sp.qual_lut = ctx->d_qual_lut; sp.int_logs = ctx->d_int_logs;
  sp.stut = (double*)ctx->d_stut.p; sp.stut_pos = (int32_t*)ctx->d_stut_pos.p;
  sp.job_t_off = (const int64_t*)m[6].p;
  sp.job_counter = (int32_t*)ctx->d_counters.p + kNumColVariants;
  // K1b in TRACE mode
  AlignParams ap;
  std::memset(&ap, 0, sizeof(ap));
  ap.pools = p.pools; ap.bases = p.bases; ap.quals = p.quals; ap.hapsides = p.hapsides; ap.hapbytes = p.hapbytes;
  ap.blocks = p.blocks; ap.reps = sp.reps; ap.progs = sp.progs; ap.rep_tabs = sp.rep_tabs;
  ap.qual_lut = ctx->d_qual_lut; ap.trans = ctx->d_trans; ap.int_logs = ctx->d_int_logs;
  ap.l_max = l_all; ap.last_scratch = (double*)ctx->d_last.p;
  ap.stut = (const double*)ctx->d_stut.p; ap.stut_pos = (const int32_t*)ctx->d_stut_pos.p;
  ap.dec = (unsigned char*)ctx->d_dec.p; ap.dec_off = p.dec_off; ap.art = (int32_t*)ctx->d_art.p; ap.art_off = p.art_off;
  ap.trace_seed_pos = (int32_t*)o[2].p;
  CU(wait_stream(ctx, s));
  ctx->trace_seconds[1] += now() - t_mark; t_mark = now();
  ctx->last_launches = 0;
  if (sp.n_jobs > 0) { CU(launch_stutter(sp, s)); ctx->last_launches++; }
  for (int v = kNumColVariants - 1; v >= 0; v--) {
    if (jobs[v].empty()) continue;
    CU(put(d.jobs[v], jobs[v], s));
    CU(put(ctx->d_job_t_off[v], job_t_off[v], s));
    ap.jobs = (const DevJob*)d.jobs[v].p; ap.n_jobs = (int32_t)jobs[v].size(); ap.n_max = n_max_v[v];
    ap.job_t_off = (const int64_t*)ctx->d_job_t_off[v].p;
    ap.job_counter = (int32_t*)ctx->d_counters.p + v;
    CU(launch_trace_forward(v, ap, HIPSTR_MAX_ALIGN_CTAS, s));
    ctx->last_launches++;
  }
  CU(launch_trace_walk(p, s));
  ctx->last_launches++;
  CU(wait_stream(ctx, s));
  ctx->trace_seconds[2] += now() - t_mark; t_mark = now();
  CU(get(ctx, out->hap_aln, p.out_aln, T * out->aln_stride));
  CU(get(ctx, out->seed_hap_pos, p.out_seed_pos, T));
  CU(get(ctx, out->stutter_size, p.out_stutter, T * HIPSTR_MAX_BLOCKS_PER_LOCUS));
  CU(get(ctx, out->span_start, p.out_span_start, T * HIPSTR_MAX_BLOCKS_PER_LOCUS));
  CU(get(ctx, out->span_len, p.out_span_len, T * HIPSTR_MAX_BLOCKS_PER_LOCUS));
  CU(get(ctx, out->flank_ins, p.out_flank_ins, T));
  CU(get(ctx, out->flank_del, p.out_flank_del, T));
  CU(get(ctx, out->n_indels, p.out_n_indels, T));
  CU(get(ctx, out->n_snps, p.out_n_snps, T));
  CU(get(ctx, out->indels, p.out_indels, T * 2 * HIPSTR_MAX_TRACE_INDELS));
  CU(get(ctx, out->snps, p.out_snps, T * 2 * HIPSTR_MAX_TRACE_SNPS));
  CU(wait_stream(ctx, s));
  ctx->trace_seconds[3] += now() - t_mark;
  end_call(ctx);
  return HIPSTR_OK;
}

}  // extern "C"
