/*
 * flatten.h -- host side of the kernel boundary: validates a hipstr_align_batch_t and
 * lowers it to the packed device layout of layout.h.
 */
#ifndef HIPSTR_B200_FLATTEN_H_
#define HIPSTR_B200_FLATTEN_H_

#include <cstring>
#include <string>
#include <functional>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "layout.h"

namespace hipstr {

/* Constant tables, computed once with glibc exactly the way the reference builds them
 * (mathops.cpp:13-21, base_quality.h:29-38, SeqAlignment/AlignmentModel.cpp:9-32) and uploaded;
 * CUDA's libm is not bit-identical to glibc's. */
struct HostTables {
  double int_logs[10000];
  double qual_lut[256][2];   /* [quality byte] -> {log P(correct), log P(error)/3} */
  double trans[3][16];       /* match->match, match->ins, match->del by homopolymer class */
  double log_one_half;
  HostTables();
};
const HostTables& host_tables();

/* Columns-per-lane variants the alignment kernel is instantiated for. */
static const int kNumColVariants = 8;
static const int kColVariants[kNumColVariants] = {2, 3, 4, 5, 6, 8, 12, 16};

/* Growable host buffer for the large staging arrays.  The allocator is pluggable so that the
 * C-ABI layer can hand out page-locked memory (cudaHostAlloc) -- H2D copies then run at full PCIe
 * speed and truly asynchronously -- while this file stays free of CUDA.  Contents are NOT
 * initialised on growth (a 130 MB zero-fill per call would cost more than the flattening). */
typedef void* (*host_alloc_fn)(size_t bytes);
typedef void (*host_free_fn)(void* p);
void set_host_allocator(host_alloc_fn alloc, host_free_fn release);
void* host_alloc(size_t bytes);
void host_free(void* p);

template <class T>
struct HostBuf {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  HostBuf() {}
  HostBuf(const HostBuf&) = delete;
  HostBuf& operator=(const HostBuf&) = delete;
  ~HostBuf() { if (p) host_free(p); }
  void resize(size_t count) {   // keeps the first min(n, count) elements
    if (count > cap) {
      size_t want = count + count / 4 + 64;
      T* q = static_cast<T*>(host_alloc(want * sizeof(T)));
      if (p) { if (n) std::memcpy(q, p, n * sizeof(T)); host_free(p); }
      p = q;
      cap = want;
    }
    n = count;
  }
  void clear() { n = 0; }
  void push_back(const T& v) { if (n == cap) { size_t old = n; resize(n + 1); n = old; } p[n++] = v; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return p; }
  const T* data() const { return p; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct FlatBatch {
  HostBuf<DevPool> pools;
  HostBuf<char> bases, quals;                /* every read padded to a multiple of 16 bytes */
  std::vector<DevHapSide> hapsides;          /* [global hap][2] */
  std::vector<uint8_t> hapbytes;
  std::vector<DevBlock> blocks;
  std::vector<DevRep> reps;
  std::vector<DevProgEntry> progs;
  std::vector<int32_t> rep_tabs;
  std::vector<uint8_t> hap_mask;             /* empty = all haplotypes */
  HostBuf<DevJob> jobs[kNumColVariants];
  int32_t n_max[kNumColVariants];            /* per variant: max read length (padded) */
  int32_t l_max[kNumColVariants];            /* per variant: max haplotype length (padded) */
  /* K1a (stutter tables): the (repeat block, allele) slots of every locus, one job per (pooled read, up to
   * HIPSTR_STUT_SLOTS_PER_JOB slots), and the slab offset of every pool inside its chunk's table buffer.  A chunk is a
   * range of pools whose tables fit the budget: K1a then K1b run chunk by chunk over the same buffer. */
  std::vector<DevSlotReps> slot_reps;
  std::vector<int32_t> locus_slot0;          /* [n_loci] first entry of slot_reps of the locus */
  HostBuf<DevStutJob> stut_jobs;
  HostBuf<int64_t> pool_t_off;
  struct Chunk {
    int32_t stut_job0, stut_job1;
    int32_t job0[kNumColVariants], job1[kNumColVariants];
    int64_t t_doubles;
  };
  std::vector<Chunk> chunks;
  int32_t stut_n_max = 16;                   /* longest read with a K1a job (padded to 16) */
  int64_t n_out = 0;
  int64_t n_alignments = 0;
  void clear();
};

/* Host threads a call may use: HIPSTR_HOST_THREADS or the hardware concurrency (at most 32), capped by the calling
 * thread's own budget when one was set (the multi-GPU driver splits the cores between its window workers). */
int host_thread_budget();
void set_host_thread_budget(int n);   /* for the calling thread; 0 = no cap */
/* Runs fn(i) for i in [0, n): the caller plus up to workers - 1 threads of one process-wide pool of parked threads
 * (created on first use, never per call -- a window of the loop issues ~50 of these).  Indices are handed out one at a
 * time; returns when every fn(i) has returned.  Safe to call from several threads at once and from inside fn. */
void parallel_run(size_t n, int workers, const std::function<void(size_t)>& fn);

/* Returns HIPSTR_OK or an error with a message. */
/* fresh_rows: give every haplotype the homopolymer classes of a from-scratch alignment (what
 * trace_optimal_aln sees: the haplotype is fixed, nothing is reused) instead of replaying the
 * reuse history of a process_reads run. */
hipstr_status_t flatten_batch(const hipstr_align_batch_t* b, FlatBatch& out, std::string& err, bool fresh_rows = false);
/* Doubles of stutter table one chunk may hold (HIPSTR_T_BUDGET_MB, default 16384 MB of the 180 GB; the device keeps two such buffers when a batch has several chunks). */
int64_t stutter_table_budget_doubles();
int64_t count_alignments(const hipstr_align_batch_t* b);

/* Per-block option index of haplotype `hap`: closed form of the reflected mixed-radix Gray
 * code walked by Haplotype::next() (SeqAlignment/Haplotype.cpp:157-196), block 0 fastest. */
void haplotype_options(int n_blocks, const int32_t* n_opts, int64_t hap, int32_t* out);

}  // namespace hipstr
#endif
