/*
 * multi_gpu.cpp -- seam B1 for a whole locus list on every GPU of the box, from C++ (hipstr_multi_*).
 *
 * The reference walks its region list serially (BamProcessor::process_regions, src/bam_processor.cpp:550-617: one
 * SeqStutterGenotyper per locus, records to the VCFWriter in region order).  Loci are independent, so here the list is cut
 * into WINDOWS of consecutive loci; a worker thread per (device, pipeline) owns a context on its device and pulls the
 * next window from a dealer as soon as it has finished one -- windows are dealt heaviest first (cost = reads of the
 * window), so rounds-per-locus differences and slow devices even out without any static split.  A window is the whole
 * path: constructor from reads -> genotype() (lockstep rounds of K5 / K1+K2+K3 / K3) -> write_vcf_record.  Several
 * pipelines per device overlap the host stages of one window (per-locus decisions, trace stitching, VCF text) with the
 * device stages of another.  The finished records are kept per locus, so the caller reads them back in locus order --
 * the order VCFWriter::add_vcf_record needs (src/vcf_writer.h:33-35) -- no matter which device produced them.
 *
 * No data-path exchange between devices: the only shared state is the dealer's counter.  When the workers of several
 * PROCESSES (one per GPU under torchrun) are to share one locus list, the caller passes its own dealer (a counter in
 * the rendezvous store); window order is a pure function of the inputs, so every process agrees on it.
 */
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "../csrc/flatten.h"
#include "seq_stutter_genotyper.h"

struct hipstr_multi {
  std::vector<int32_t> devices;
  int32_t pipelines = 1;
  std::vector<hipstr_ctx_t*> ctxs;            // [device][pipeline]
  // the last run
  std::vector<std::string> records;           // per locus, empty = none
  std::vector<int32_t> pos;
  std::vector<int32_t> window_owner;          // worker that processed the window, -1 = another process
  std::vector<int32_t> windows_done;          // per worker
  std::vector<double> busy_seconds;           // per worker
  double seconds[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int64_t n_alignments = 0, n_traces = 0, h2d_bytes = 0, d2h_bytes = 0, gpu_launches = 0;
  std::string last_error;
};

namespace {

/* windows in dealing order: heaviest (most reads) first, ties by position */
std::vector<int32_t> window_order(int32_t n_loci, int32_t window_loci, const int32_t* locus_read_off) {
  const int32_t n_windows = (n_loci + window_loci - 1) / window_loci;
  std::vector<int32_t> order(n_windows);
  std::iota(order.begin(), order.end(), 0);
  auto cost = [&](int32_t w) {
    const int32_t l0 = w * window_loci, l1 = std::min(n_loci, l0 + window_loci);
    return (int64_t)locus_read_off[l1] - locus_read_off[l0];
  };
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost(a) > cost(b); });
  return order;
}

}  // namespace

extern "C" {

hipstr_status_t hipstr_multi_create(int32_t n_devices, const int32_t* devices, int32_t pipelines_per_device, hipstr_multi_t** out) {
  if (n_devices <= 0 || !devices || pipelines_per_device <= 0 || pipelines_per_device > 8 || !out) return HIPSTR_ERR_BAD_ARG;
  *out = nullptr;
  hipstr_multi* m = new hipstr_multi();
  m->devices.assign(devices, devices + n_devices);
  m->pipelines = pipelines_per_device;
  for (int d = 0; d < n_devices; d++)
    for (int p = 0; p < pipelines_per_device; p++) {
      hipstr_ctx_t* ctx = nullptr;
      const hipstr_status_t st = hipstr_create(devices[d], &ctx);
      if (st != HIPSTR_OK) {
        for (hipstr_ctx_t* c : m->ctxs) hipstr_destroy(c);
        delete m;
        return st;
      }
      m->ctxs.push_back(ctx);
    }
  *out = m;
  return HIPSTR_OK;
}

void hipstr_multi_destroy(hipstr_multi_t* m) {
  if (!m) return;
  for (hipstr_ctx_t* c : m->ctxs) hipstr_destroy(c);
  delete m;
}

const char* hipstr_multi_last_error(const hipstr_multi_t* m) { return m ? m->last_error.c_str() : "null handle"; }

int32_t hipstr_multi_num_windows(int32_t n_loci, int32_t window_loci) {
  return window_loci > 0 ? (n_loci + window_loci - 1) / window_loci : -1;
}

hipstr_status_t hipstr_multi_window_order(int32_t n_loci, int32_t window_loci, const int32_t* locus_read_off, int32_t* order) {
  if (n_loci < 0 || window_loci <= 0 || !locus_read_off || !order) return HIPSTR_ERR_BAD_ARG;
  const std::vector<int32_t> o = window_order(n_loci, window_loci, locus_read_off);
  std::copy(o.begin(), o.end(), order);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_multi_genotype(hipstr_multi_t* m, int32_t n_loci, const int32_t* region_start, const int32_t* region_stop,
                                      const int32_t* period, const char* const* chrom_seq, const double* stutter,
                                      const hipstr_locus_reads_t* reads, const hipstr_vcf_loci_t* vcf_loci,
                                      const hipstr_vcf_options_t* vcf_options, int32_t max_total_haplotypes,
                                      int32_t max_flank_haplotypes, double min_flank_freq, int32_t window_loci,
                                      hipstr_next_window_fn next_window, void* next_window_user, uint8_t* locus_ok) {
  if (!m || n_loci < 0 || !region_start || !region_stop || !period || !chrom_seq || !stutter || !reads || !vcf_loci || window_loci <= 0)
    return HIPSTR_ERR_BAD_ARG;
  hipstr_vcf_options_t def;
  hipstr_vcf_default_options(&def);
  const hipstr_vcf_options_t* opt = vcf_options ? vcf_options : &def;
  const int32_t n_windows = hipstr_multi_num_windows(n_loci, window_loci);
  const std::vector<int32_t> order = window_order(n_loci, window_loci, reads->locus_read_off);
  const int n_workers = (int)m->ctxs.size();
  m->records.assign((size_t)n_loci, std::string());
  m->pos.assign((size_t)n_loci, 0);
  m->window_owner.assign((size_t)n_windows, -1);
  m->windows_done.assign((size_t)n_workers, 0);
  m->busy_seconds.assign((size_t)n_workers, 0.0);
  std::fill(m->seconds, m->seconds + 9, 0.0);
  m->n_alignments = m->n_traces = m->h2d_bytes = m->d2h_bytes = m->gpu_launches = 0;
  if (locus_ok) std::memset(locus_ok, 0, (size_t)n_loci);
  std::atomic<int32_t> counter(0);
  std::atomic<int> failed(0);
  std::mutex mu;   // stats + error text
  const int threads_per_worker = std::max(1, hipstr::host_thread_budget() / n_workers);
  auto worker = [&](int w) {
    hipstr::set_host_thread_budget(threads_per_worker);
    hipstr_ctx_t* ctx = m->ctxs[(size_t)w];
    for (;;) {
      if (failed.load()) return;
      const int32_t k = next_window ? next_window(next_window_user) : counter.fetch_add(1);
      if (k < 0 || k >= n_windows) return;
      const double t0 = hipstr::now_s();
      const int32_t win = order[(size_t)k];
      const int32_t l0 = win * window_loci, l1 = std::min(n_loci, l0 + window_loci), L = l1 - l0;
      // the window's view of the per-locus arrays; read-level arrays stay absolute (locus_read_off holds absolute indices)
      hipstr_locus_reads_t rd = *reads;
      rd.locus_read_off = reads->locus_read_off + l0;
      rd.locus_sample_off = reads->locus_sample_off + l0;
      rd.haploid = reads->haploid ? reads->haploid + l0 : nullptr;
      hipstr_vcf_loci_t vl = *vcf_loci;
      vl.chrom = vcf_loci->chrom + l0;
      vl.name = vcf_loci->name ? vcf_loci->name + l0 : nullptr;
      vl.region_start = vcf_loci->region_start + l0;
      vl.region_stop = vcf_loci->region_stop + l0;
      vl.period = vcf_loci->period + l0;
      vl.chrom_seq = vcf_loci->chrom_seq + l0;
      vl.locus_sample_names = vcf_loci->locus_sample_names + (reads->locus_sample_off[l0] - reads->locus_sample_off[0]);
      hipstr_genotyper_t* g = nullptr;
      std::vector<uint8_t> ok((size_t)L, 0);
      hipstr_status_t st = hipstr_genotyper_create_from_reads(ctx, L, region_start + l0, region_stop + l0, period + l0, chrom_seq + l0,
                                                              stutter + 6 * (size_t)l0, &rd, &g);
      const char* what = "hipstr_genotyper_create_from_reads";
      if (st == HIPSTR_OK) { st = hipstr_genotyper_genotype(g, max_total_haplotypes, max_flank_haplotypes, min_flank_freq, 1, ok.data()); what = "hipstr_genotyper_genotype"; }
      if (st == HIPSTR_OK) { st = hipstr_genotyper_write_vcf(g, &vl, opt); what = "hipstr_genotyper_write_vcf"; }
      if (st != HIPSTR_OK) {
        std::lock_guard<std::mutex> lock(mu);
        m->last_error = std::string(what) + ": " + (g ? hipstr_genotyper_last_error(g) : "construction failed");
        failed.store((int)st);
        if (g) hipstr_genotyper_destroy(g);
        return;
      }
      for (int32_t l = 0; l < L; l++) {
        m->records[(size_t)(l0 + l)] = g->batch.loci[(size_t)l].vcf_record_;
        m->pos[(size_t)(l0 + l)] = g->batch.loci[(size_t)l].vcf_pos_;
        if (locus_ok) locus_ok[l0 + l] = ok[(size_t)l];
      }
      {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < 9; i++) m->seconds[i] += g->batch.seconds[i];
        m->n_alignments += g->batch.n_alignments;
        m->n_traces += g->batch.n_traces;
        m->h2d_bytes += g->batch.h2d_bytes; m->d2h_bytes += g->batch.d2h_bytes; m->gpu_launches += g->batch.gpu_launches;
        m->window_owner[(size_t)win] = w;
        m->windows_done[(size_t)w]++;
        m->busy_seconds[(size_t)w] += hipstr::now_s() - t0;
      }
      hipstr_genotyper_destroy(g);
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < n_workers; w++) pool.emplace_back(worker, w);
  worker(0);
  for (std::thread& t : pool) t.join();
  hipstr::set_host_thread_budget(0);
  return failed.load() ? (hipstr_status_t)failed.load() : HIPSTR_OK;
}

int32_t hipstr_multi_locus_record(const hipstr_multi_t* m, int32_t locus, int32_t* pos, char* out, int32_t cap) {
  if (!m || locus < 0 || locus >= (int32_t)m->records.size()) return -1;
  const std::string& r = m->records[(size_t)locus];
  if (pos) *pos = m->pos[(size_t)locus];
  if ((int32_t)r.size() + 1 > cap || !out) return -((int32_t)r.size() + 1);
  std::memcpy(out, r.c_str(), r.size() + 1);
  return (int32_t)r.size();
}

hipstr_status_t hipstr_multi_emit_records(const hipstr_multi_t* m, const hipstr_vcf_loci_t* loci, hipstr_vcf_writer_t* w) {
  if (!m || !loci || !w) return HIPSTR_ERR_BAD_ARG;
  for (size_t l = 0; l < m->records.size(); l++) {
    if (m->records[l].empty()) continue;
    const hipstr_status_t st = hipstr_vcf_writer_add_record(w, loci->chrom[l], m->pos[l], m->records[l].c_str());
    if (st != HIPSTR_OK) return st;
  }
  return HIPSTR_OK;
}

hipstr_status_t hipstr_multi_stats(const hipstr_multi_t* m, int64_t* n_alignments, int64_t* n_traces, double* seconds9,
                                   int32_t* windows_per_worker, double* busy_seconds_per_worker) {
  if (!m) return HIPSTR_ERR_BAD_ARG;
  if (n_alignments) *n_alignments = m->n_alignments;
  if (n_traces) *n_traces = m->n_traces;
  if (seconds9) std::copy(m->seconds, m->seconds + 9, seconds9);
  if (windows_per_worker) std::copy(m->windows_done.begin(), m->windows_done.end(), windows_per_worker);
  if (busy_seconds_per_worker) std::copy(m->busy_seconds.begin(), m->busy_seconds.end(), busy_seconds_per_worker);
  return HIPSTR_OK;
}

hipstr_status_t hipstr_multi_traffic(const hipstr_multi_t* m, int64_t* h2d_bytes, int64_t* d2h_bytes, int64_t* gpu_launches) {
  if (!m) return HIPSTR_ERR_BAD_ARG;
  if (h2d_bytes) *h2d_bytes = m->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = m->d2h_bytes;
  if (gpu_launches) *gpu_launches = m->gpu_launches;
  return HIPSTR_OK;
}

int32_t hipstr_multi_num_workers(const hipstr_multi_t* m) { return m ? (int32_t)m->ctxs.size() : -1; }

}  // extern "C"
