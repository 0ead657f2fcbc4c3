/*
 * hostsim.cpp -- TEST INFRASTRUCTURE, never part of the product library.
 *
 * The host side of the hot path (the lockstep genotyping loop, VCF records, left alignment, the region driver) calls
 * the GPU only through the device entry points of include/hipstr_b200.h.  This file provides those entry points on top
 * of the CPU oracle (oracle/liboracle.so) so that the HOST LOGIC can be exercised and profiled in a container without
 * a GPU (`-m "not gpu"` tests, gprof).  It is linked only into tests/hostsim/libhipstr_hostsim.so; the product
 * (hipstr_b200/libhipstr_b200.so) has no such path and fails with HIPSTR_ERR_NO_DEVICE without a GPU.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "../../hipstr_b200/csrc/flatten.h"
#include "../../oracle/hipstr_oracle.h"

struct hipstr_ctx { std::string last_error; double trace_seconds[4] = {0, 0, 0, 0}; };

/* Record / replay of the simulated device calls, for profiling the HOST side alone: with HIPSTR_SIM_RECORD=<file> every
 * call appends its outputs to the file; with HIPSTR_SIM_REPLAY=<file> the same sequence of calls (same inputs, ONE host
 * thread, one window at a time) gets the recorded outputs back without running the oracle; the tape rewinds at its end. */
namespace {
struct Tape {
  FILE* f = nullptr;
  bool replay = false;
  Tape() {
    if (const char* p = std::getenv("HIPSTR_SIM_REPLAY")) { f = std::fopen(p, "rb"); replay = true; }
    else if (const char* q = std::getenv("HIPSTR_SIM_RECORD")) f = std::fopen(q, "wb");
  }
  bool playing() {
    if (!(f && replay)) return false;
    const int c = std::fgetc(f);
    if (c == EOF) std::rewind(f); else std::ungetc(c, f);
    return true;
  }
  template <class T> void io(T* p, size_t n) {
    if (!f || !p || n == 0) return;
    if (replay) { if (std::fread(p, sizeof(T), n, f) != n) { std::fprintf(stderr, "hostsim: tape does not match the calls\n"); std::abort(); } }
    else std::fwrite(p, sizeof(T), n, f);
  }
};
Tape& tape() { static Tape t; return t; }
}  // namespace

extern "C" {

const char* hipstr_version(void) { return "hipstr_b200 host simulation over the CPU oracle (tests only)"; }
hipstr_status_t hipstr_create(int, hipstr_ctx_t** out) { if (!out) return HIPSTR_ERR_BAD_ARG; *out = new hipstr_ctx(); return HIPSTR_OK; }
void hipstr_destroy(hipstr_ctx_t* c) { delete c; }
const char* hipstr_last_error(const hipstr_ctx_t* c) { return c ? c->last_error.c_str() : "no context"; }
hipstr_status_t hipstr_set_stream(hipstr_ctx_t*, void*) { return HIPSTR_OK; }
int32_t hipstr_last_launch_count(const hipstr_ctx_t*) { return 0; }
int64_t hipstr_batch_num_alignments(const hipstr_align_batch_t* b) { return b ? hipstr::count_alignments(b) : 0; }
hipstr_status_t hipstr_enable_timing(hipstr_ctx_t*, int) { return HIPSTR_OK; }
float hipstr_last_kernel_ms(const hipstr_ctx_t*) { return 0.f; }
void hipstr_trace_seconds(const hipstr_ctx_t*, double* s) { if (s) for (int i = 0; i < 4; i++) s[i] = 0; }
void hipstr_last_traffic(const hipstr_ctx_t*, int64_t* a, int64_t* b, int32_t* n) { if (a) *a = 0; if (b) *b = 0; if (n) *n = 0; }
hipstr_status_t hipstr_collect_timing(hipstr_ctx_t*, double* a, double* b, int32_t* n) { if (a) *a = 0; if (b) *b = 0; if (n) *n = 0; return HIPSTR_OK; }
hipstr_status_t hipstr_debug_lastcols(hipstr_ctx_t*, double*, int32_t) { return HIPSTR_ERR_UNSUPPORTED; }

/* resident (device-pointer) entry points have no meaning without a device */
hipstr_status_t hipstr_upload_batch(hipstr_ctx_t*, const hipstr_align_batch_t*, hipstr_dev_batch_t**) { return HIPSTR_ERR_UNSUPPORTED; }
hipstr_status_t hipstr_align_batch_dev(hipstr_ctx_t*, const hipstr_dev_batch_t*, double*, int32_t*) { return HIPSTR_ERR_UNSUPPORTED; }
void hipstr_free_batch(hipstr_ctx_t*, hipstr_dev_batch_t*) {}
hipstr_status_t hipstr_upload_genotype_batch(hipstr_ctx_t*, const hipstr_align_batch_t*, const hipstr_reads_batch_t*, hipstr_dev_genotype_t**) { return HIPSTR_ERR_UNSUPPORTED; }
hipstr_status_t hipstr_genotype_batch_dev(hipstr_ctx_t*, const hipstr_dev_genotype_t*, const hipstr_genotype_out_t*) { return HIPSTR_ERR_UNSUPPORTED; }
void hipstr_free_genotype_batch(hipstr_ctx_t*, hipstr_dev_genotype_t*) {}

hipstr_status_t hipstr_align_batch_host(hipstr_ctx_t* c, const hipstr_align_batch_t* b, double* ll, int32_t* pos) {
  if (!c || !b || !ll) return HIPSTR_ERR_BAD_ARG;
  return oracle_align_batch(b, ll, pos) == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_scatter_pool_lls_host(hipstr_ctx_t*, int32_t R, int32_t H, const double* pool_ll, const int32_t* pool_seed,
                                             const int32_t* pool_index, const uint8_t* second_mate, const uint8_t* copy_read,
                                             const uint8_t* realign_hap, double* read_ll, int32_t* read_seed) {
  return oracle_scatter_pool_lls(R, H, pool_ll, pool_seed, pool_index, second_mate, copy_read, realign_hap, read_ll, read_seed) == 0
             ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_posteriors_host(hipstr_ctx_t*, int32_t n_loci, const int32_t* lro, const int32_t* lso, const int32_t* n_haps,
                                       const uint8_t* haploid, const double* read_ll, const double* p1, const double* p2,
                                       const int32_t* label, const int32_t* weight, double* post, double* sll, int32_t* best,
                                       double* tot) {
  size_t n_post = 0;
  for (int l = 0; l < n_loci; l++) n_post += (size_t)(lso[l + 1] - lso[l]) * n_haps[l] * n_haps[l];
  const size_t S = n_loci ? (size_t)lso[n_loci] : 0;
  auto io = [&] { tape().io(post, n_post); tape().io(sll, S); tape().io(best, 2 * S); tape().io(tot, (size_t)n_loci); };
  if (tape().playing()) { io(); return HIPSTR_OK; }
  const int rc = oracle_posteriors(n_loci, lro, lso, n_haps, haploid, read_ll, p1, p2, label, weight, post, sll, best, tot);
  io();
  return rc == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
/* K1 + K2 + K3 of a batch (seq_stutter_genotyper.cpp:519-568,638-639), masks with the in-place semantics of the product */
hipstr_status_t hipstr_genotype_batch_host(hipstr_ctx_t* c, const hipstr_align_batch_t* b, const hipstr_reads_batch_t* r,
                                           const hipstr_genotype_out_t* o) {
  if (!c || !b || !r || !o || !o->read_ll || !o->post || !o->sample_ll) return HIPSTR_ERR_BAD_ARG;
  const int L = b->n_loci;
  size_t n_ll = 0, n_post = 0;
  for (int l = 0; l < L; l++) {
    const size_t H = (size_t)(b->locus_hap_off[l + 1] - b->locus_hap_off[l]);
    n_ll += (size_t)(r->locus_read_off[l + 1] - r->locus_read_off[l]) * H;
    n_post += (size_t)(r->locus_sample_off[l + 1] - r->locus_sample_off[l]) * H * H;
  }
  const size_t n_r = L ? (size_t)r->locus_read_off[L] : 0, n_s = L ? (size_t)r->locus_sample_off[L] : 0;
  auto io = [&] {
    tape().io(o->read_ll, n_ll); tape().io(o->read_seed, n_r); tape().io(o->post, n_post); tape().io(o->sample_ll, n_s);
    tape().io(o->best, 2 * n_s); tape().io(o->total_ll, (size_t)L);
  };
  if (tape().playing()) { io(); return HIPSTR_OK; }
  std::vector<double> pool_ll((size_t)(L ? b->locus_out_off[L] : 0), 0.0);
  if (oracle_align_batch(b, pool_ll.data(), nullptr) != 0) { c->last_error = "oracle_align_batch failed"; return HIPSTR_ERR_BAD_ARG; }
  std::vector<int32_t> n_haps((size_t)L);
  int64_t ll_at = 0;
  for (int l = 0; l < L; l++) {
    const int H = (int)(b->locus_hap_off[l + 1] - b->locus_hap_off[l]);
    n_haps[l] = H;
    const int r0 = r->locus_read_off[l], R = r->locus_read_off[l + 1] - r0, p0 = b->locus_pool_off[l];
    /* pools the mask excludes keep their old values: scatter only copies reads with copy_read set */
    std::vector<uint8_t> copy((size_t)R, 1);
    for (int i = 0; i < R; i++) {
      const bool pool_on = !b->realign_pool || b->realign_pool[p0 + r->pool_index[r0 + i]];
      copy[i] = (r->copy_read ? r->copy_read[r0 + i] != 0 : true) && pool_on;
    }
    if (oracle_scatter_pool_lls(R, H, pool_ll.data() + b->locus_out_off[l], b->pool_seed + p0, r->pool_index + r0, r->second_mate + r0,
                                copy.data(), b->realign_hap ? b->realign_hap + b->locus_hap_off[l] : nullptr, o->read_ll + ll_at,
                                o->read_seed ? o->read_seed + r0 : nullptr) != 0)
      return HIPSTR_ERR_BAD_ARG;
    ll_at += (int64_t)R * H;
  }
  const int rc = oracle_posteriors(L, r->locus_read_off, r->locus_sample_off, n_haps.data(), r->haploid, o->read_ll, r->log_p1, r->log_p2,
                                   r->sample_label, r->read_weight, o->post, o->sample_ll, o->best, o->total_ll);
  io();
  return rc == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_extract_genotypes_host(hipstr_ctx_t*, int32_t n_loci, const int32_t* lso, const int32_t* n_haps,
                                              const int32_t* n_variants, const int32_t* h2a, const uint8_t* haploid, const double* post,
                                              const double* sll, int32_t* best_hap, int32_t* best_gt, double* lp, double* lu, double* hlp,
                                              double* hlu, double* gl, double* pgl, double* gl_diff, int32_t* pl) {
  size_t n_gl = 0, n_pgl = 0;
  for (int l = 0; l < n_loci; l++) {
    const size_t S = (size_t)(lso[l + 1] - lso[l]), V = (size_t)n_variants[l];
    n_gl += S * (haploid[l] ? V : V * (V + 1) / 2);
    n_pgl += S * (haploid[l] ? V : V * V);
  }
  const size_t S = n_loci ? (size_t)lso[n_loci] : 0;
  auto io = [&] {
    tape().io(best_hap, 2 * S); tape().io(best_gt, 2 * S); tape().io(lp, S); tape().io(lu, S); tape().io(hlp, S); tape().io(hlu, S);
    tape().io(gl, n_gl); tape().io(pgl, n_pgl); tape().io(gl_diff, S); tape().io(pl, n_gl);
  };
  if (tape().playing()) { io(); return HIPSTR_OK; }
  const int rc = oracle_extract_genotypes(n_loci, lso, n_haps, n_variants, h2a, haploid, post, sll, best_hap, best_gt, lp, lu, hlp, hlu, gl,
                                          pgl, gl_diff, pl);
  io();
  return rc == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_trace_batch_host(hipstr_ctx_t*, const hipstr_align_batch_t* b, const int32_t* block_start, int32_t n,
                                        const int32_t* tp, const int32_t* th, const hipstr_trace_out_t* out) {
  if (n == 0) return HIPSTR_OK;
  const size_t T = (size_t)n;
  auto io = [&] {
    tape().io(out->hap_aln, T * (size_t)out->aln_stride); tape().io(out->seed_hap_pos, T); tape().io(out->stutter_size, 8 * T);
    tape().io(out->span_start, 8 * T); tape().io(out->span_len, 8 * T); tape().io(out->flank_ins, T); tape().io(out->flank_del, T);
    tape().io(out->n_indels, T); tape().io(out->indels, T * 2 * HIPSTR_MAX_TRACE_INDELS); tape().io(out->n_snps, T);
    tape().io(out->snps, T * 2 * HIPSTR_MAX_TRACE_SNPS);
  };
  if (tape().playing()) { io(); return HIPSTR_OK; }
  const int rc = oracle_trace_batch(b, block_start, n, tp, th, out);
  io();
  return rc == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_em_train_host(hipstr_ctx_t*, const hipstr_em_batch_t* b, int32_t max_iter, double a, double f, double* params,
                                     uint8_t* conv, int32_t* iters, double* ll) {
  return oracle_em_train(b, max_iter, a, f, params, conv, iters, ll) == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}
hipstr_status_t hipstr_nw_align_batch_host(hipstr_ctx_t*, int32_t n_pairs, const int32_t* ref_off, const char* ref_seqs,
                                           const int32_t* read_off, const char* read_seqs, int32_t end_penalty, int32_t ops_stride,
                                           char* ops, int32_t* ops_len, float* score) {
  for (int i = 0; i < n_pairs; i++) {
    float sc = 0;
    char* o = ops + (size_t)i * ops_stride;
    if (oracle_nw_align(ref_seqs + ref_off[i], ref_off[i + 1] - ref_off[i], read_seqs + read_off[i], read_off[i + 1] - read_off[i],
                        end_penalty, o, &sc) != 0) return HIPSTR_ERR_BAD_ARG;
    if (ops_len) ops_len[i] = (int32_t)std::strlen(o);
    if (score) score[i] = sc;
  }
  return HIPSTR_OK;
}
hipstr_status_t hipstr_snp_phasing_batch_host(hipstr_ctx_t*, const hipstr_snp_phasing_t* b, double* p1, double* p2, int32_t* counts) {
  return oracle_snp_phasing(b, p1, p2, counts) == 0 ? HIPSTR_OK : HIPSTR_ERR_BAD_ARG;
}

}  // extern "C"

/* Test hook for the product's host thread pool (hipstr::parallel_run, csrc/flatten.cpp): `callers` threads call it at the same
 * time, each `rounds` times over `n` indices with `workers` workers, and every index calls it again from inside (nested use,
 * as the loop does when a per-locus step lowers a batch).  Returns the number of wrong sums (0 = every index ran exactly once). */
#include <atomic>
#include <thread>
extern "C" int64_t hostsim_exercise_thread_pool(int32_t callers, int32_t rounds, int32_t n, int32_t workers) {
  std::atomic<int64_t> wrong(0);
  auto one_caller = [&](int c) {
    for (int r = 0; r < rounds; r++) {
      std::vector<std::atomic<int32_t> > hits((size_t)n);
      for (auto& h : hits) h = 0;
      std::atomic<int64_t> inner_total(0);
      hipstr::parallel_run((size_t)n, workers, [&](size_t i) {
        hits[i]++;
        std::atomic<int64_t> inner(0);
        hipstr::parallel_run(4, 2, [&](size_t k) { inner += (int64_t)(k + 1); });   // 1 + 2 + 3 + 4
        inner_total += inner.load();
      });
      for (auto& h : hits) wrong += h.load() != 1;
      wrong += inner_total.load() != 10 * (int64_t)n;
      (void)c;
    }
  };
  std::vector<std::thread> threads;
  for (int c = 0; c < callers; c++) threads.emplace_back(one_caller, c);
  for (auto& t : threads) t.join();
  return wrong.load();
}

