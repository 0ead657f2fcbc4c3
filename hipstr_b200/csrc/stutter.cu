/*
 * stutter.cu -- K1a: the stutter tables of the read x haplotype HMM.
 *
 * Inside a repeat block the reference evaluates, for every read column j and each of the 13 PCR-artifact sizes D,
 *     log_prob_pcr_artifact(D) + align_stutter_region_reverse(read prefix ending at j, D) + M[row above the block][j - base_len]
 * (SeqAlignment/HapAligner.cpp:76-100 over StutterAlignerClass.cpp:12-162).  The first two terms depend only on the
 * READ and the repeat ALLELE -- not on the flanks, not on the DP state -- so they are computed here, once per
 * (pooled read, allele), as a pure map over read columns, and written to the stutter tables
 *     T[pool][slot][artifact][column]
 * that K1b (kernels.cu) folds into the DP with the third term.  Splitting the map out of the wavefront kernel
 *   - gives every lane a whole column to itself (32 of 32 lanes busy; the two read sides never share a warp, so all
 *     lanes of a warp replay the same host-unrolled position walk in lock step),
 *   - lets a table be shared by every haplotype that carries the allele (flank variants after assembly),
 *   - and needs half K1b's registers, so twice the warps hide the FP64 / shared-memory latencies.
 * The tables go through HBM (24 KB per alignment written + read): the memory system is otherwise idle in this path.
 *
 * One CTA of 128 threads per job = (pooled read, up to 8 allele slots of its locus).  The read's packed bases and
 * qualities arrive by TMA bulk copy (double-buffered: the copy of the NEXT job's read is in flight while this one is
 * processed) and are expanded to the emission table val[column][5] in shared memory, columns in SIDE order (left of
 * the seed forwards, right of the seed reversed; HapAligner.cpp:579-585,606-609).  Per slot:
 *   pass 1  thread per column: match_probs_[q] with snapshots at k*period terms = del_probs_ (load_read, :12-53);
 *   pass 2  thread per column: the 6 deletion and 6 insertion walks (align_pcr_deletion_reverse :106-150,
 *           align_pcr_insertion_reverse :59-104) replayed from the host-unrolled programs (layout.h DevProgEntry),
 *           each finished with the reference's approximate log-sum-exp (mathops.cpp:97-106).
 * Every double operation is done in the reference's order; the float pieces are the replicas of fastapprox.cuh.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hipstr_b200.h"
#include "fastapprox.cuh"
#include "kcommon.cuh"
#include "kernels.h"
#include "layout.h"

namespace hipstr {

#define STUT_THREADS 128
#define STUT_WARPS (STUT_THREADS / 32)
#ifndef STUT_TERM_SLOTS
#define STUT_TERM_SLOTS 16                      /* terms of one walk kept in shared memory per lane */
#endif
#define STUT_TERM_STRIDE 256                    /* bytes between two slots of one lane (32 lanes x 8) */

struct StutCtx {
  const DevProgEntry* progs;
  const int32_t* diag;       // byte offsets of the right-anchored diagonal (DevRep::diag_off)
  const int32_t* ins_tab;    // byte offsets of the periodic-copy sum (DevRep::ins_off)
  const DevRep* rep;
  const double* int_logs;
  unsigned val;              // shared address of the side's column 0 of the emission table
  const uint8_t* code;       // shared: base codes of the side's columns
  const double* match;       // shared (this warp's): match_probs_ by side column
  unsigned terms;            // shared address of this lane's term slot 0
  int B, p, n_side;
};

// One step of a walk (layout.h DevProgEntry as int4: x = pos, y = moves, then z / w = the two offsets or the double):
// either lp = (lp - val[col + off_a]) + val[col + off_b], `units` times for insertions, and the term is lp itself, or
// the term is lp + logrun.  `moves` is the same for the whole warp.  The term goes to its shared-memory slot (when the
// walk still has one) and into the running maximum; both arms do that themselves so that neither copies a double.
template <bool INS>
__device__ __forceinline__ void stut_step(double& lp, double& mx, unsigned col, const int4& e, int units, int stride, bool keep,
                                          unsigned slot) {
  if (e.y) {
    if (INS) {
      unsigned a = col + e.z, b = col + e.w;
#pragma unroll 2
      for (int m = 0; m < units; m++, a -= stride, b -= stride) {
        lp -= lds_f64(a);
        lp += lds_f64(b);
      }
    } else {
      lp -= lds_f64(col + e.z);
      lp += lds_f64(col + e.w);
    }
    if (keep) sts_f64(slot, lp);
    mx = dmax(mx, lp);
  } else {
    const double term = lp + __hiloint2double(e.w, e.z);
    if (keep) sts_f64(slot, term);
    mx = dmax(mx, term);
  }
}

// Replays one position walk for read column j and returns fast_log_sum_exp of its terms (mathops.cpp:97-106).
// The whole warp is on the same (side, allele, artifact size): program entries, the `moves` branch and the loop
// bound are uniform; a lane only differs in where its read prefix ends (`stop`).  Terms wait in shared memory for the
// maximum; a walk with more terms than slots is replayed for the ones that did not fit.
// TRACE also tracks the reference's best artifact position (StutterAlignerClass.cpp:92-95,137-140: the position counter
// after a step is one above the next step's position, so best_pos = 1 - i == -next.pos).
template <bool INS, bool TRACE>
__device__ __forceinline__ double stut_walk(const StutCtx& c, int prog_index, int stop, int j, int units, double lp0,
                                            int tail_base, int& best_pos) {
  const int4* prog = reinterpret_cast<const int4*>(c.progs + prog_index);
  const unsigned col = c.val + (INS ? (j - c.p) : j) * HIPSTR_COL_BYTES;
  const int stride = c.p * HIPSTR_COL_BYTES;
  const int warp_stop = __reduce_min_sync(FULL, stop);
  double lp = lp0, mx = lp0, best = lp0;
  best_pos = 0;
  const bool left_align = TRACE && __ldg(&c.rep->left_align) != 0;
  sts_f64(c.terms, lp0);
  int cnt = 0, s = 0;
  unsigned slot = c.terms;
  const int4* pe = prog;
  // two entries per trip: no register shuffling for the look-ahead, half the loop overhead
  for (;;) {
    const int4 e0 = __ldg(pe), e1 = __ldg(pe + 1);
    if (e0.x <= warp_stop) break;
    if (e0.x > stop) {
      stut_step<INS>(lp, mx, col, e0, units, stride, s + 1 < STUT_TERM_SLOTS, slot + STUT_TERM_STRIDE);
      cnt++;
      if (TRACE && (lp > best || (left_align && lp == best))) { best_pos = -e1.x; best = lp; }
    }
    if (e1.x <= warp_stop) { s += 1; break; }
    if (e1.x > stop) {
      stut_step<INS>(lp, mx, col, e1, units, stride, s + 2 < STUT_TERM_SLOTS, slot + 2 * STUT_TERM_STRIDE);
      cnt++;
      if (TRACE && (lp > best || (left_align && lp == best))) { best_pos = -__ldg(&pe[2].x); best = lp; }
    }
    s += 2;
    pe += 2;
    slot += 2 * STUT_TERM_STRIDE;
  }
  // the entry this lane stopped at tells how many artifact positions are left (they all share lp)
  const int fin = __ldg(&prog[cnt].x);
  const bool has_tail = INS ? (fin > -tail_base) : (-fin < tail_base);
  double tail = 0.0;
  if (has_tail) {
    tail = __ldg(c.int_logs + (tail_base + fin)) + lp;
    mx = dmax(mx, tail);
  }
  double total = has_tail ? lse_term_near(tail, mx) : 0.0;
  const int n = cnt + 1;                                  // cached terms of this lane (slot 0 = lp0)
  const int n_warp = min(s + 1, STUT_TERM_SLOTS);
  slot = c.terms;
#pragma unroll 4
  for (int t = 0; t < n_warp; t++, slot += STUT_TERM_STRIDE)     // branch-free: a slot this lane did not fill adds 0
    total += lse_term_masked(lds_f64(slot), mx, t < n);
  if (n > STUT_TERM_SLOTS) {   // rare: a walk longer than the cache -> replay it for the terms that did not fit
    lp = lp0;
    for (int t = 0; t < cnt; t++) {
      const int4 e = __ldg(prog + t);
      double term = lp + __hiloint2double(e.w, e.z), unused = 0.0;
      if (e.y) { stut_step<INS>(lp, unused, col, e, units, stride, false, 0); term = lp; }
      if (t + 1 >= STUT_TERM_SLOTS) total += lse_term_near(term, mx);
    }
  }
  return lse_finish(mx, total);
}
// The 13 table entries of one read column (HapAligner.cpp:76-100 without pre_prob).  On entry the deletion rows of
// the table hold what pass 1 left there: match_probs_[q] - del_probs_[q][k-1] for q = j + k*period inside the read, or
// the whole first term (prior included) where the read ends before q.
template <bool TRACE>
__device__ __forceinline__ void stut_column(const StutCtx& c, int j, bool live, double* tcol, int32_t* pcol, int pitch) {
  const int B = c.B, p = c.p;
  const DevRep* rep = c.rep;
  const unsigned colj = c.val + j * HIPSTR_COL_BYTES;
#pragma unroll 1
  for (int k = HIPSTR_MAX_ARTIFACT_UNITS; k >= 1; k--) {   // deletions of k units
    const int D = -k * p;
    if (B + D < 0) continue;                               // impossible size: K1b never reads the entry
    const int base_len = min(B + D, j + 1);
    double* cell = tcol + (HIPSTR_MAX_ARTIFACT_UNITS - k) * pitch;
    const double v = *cell;
    const double lp0 = (j - D <= c.n_side - 1) ? -__ldg(c.int_logs + (B + D + 1)) + v : v;
    int pos;
    const double pr = stut_walk<false, TRACE>(c, __ldg(rep->prog_off + k), -base_len, j, k, lp0, B + D, pos);
    if (live) *cell = __ldg(rep->art + (HIPSTR_MAX_ARTIFACT_UNITS - k)) + pr;
    if (TRACE && live) pcol[(HIPSTR_MAX_ARTIFACT_UNITS - k) * pitch] = pos;
  }
  if (live) tcol[HIPSTR_MAX_ARTIFACT_UNITS * pitch] = __ldg(rep->art + HIPSTR_MAX_ARTIFACT_UNITS) + c.match[j];
  if (TRACE && live) pcol[HIPSTR_MAX_ARTIFACT_UNITS * pitch] = -1;
  double ins_acc = 0.0;    // ins_probs_ running sum (StutterAlignerClass.cpp:38-51)
  int ins_t = 0;
  const double ins_prior = -__ldg(c.int_logs + (B + 1));
  const int ins_prog = __ldg(rep->prog_off);
#pragma unroll 1
  for (int k = 1; k <= HIPSTR_MAX_ARTIFACT_UNITS; k++) {    // insertions of k units
    const int D = k * p;
    const int base_len = min(B + D, j + 1);
    const int upto = min(D, j + 1);   // at most j+1 read bases exist
    for (; ins_t < upto; ins_t++) {
      const int off = __ldg(c.ins_tab + ins_t);
      ins_acc += lds_f64(off != -1 ? colj + off : colj - ins_t * HIPSTR_COL_BYTES + 8 * c.code[j - ins_t]);
    }
    double lp0 = ins_prior + ins_acc;
    lp0 += (base_len > D) ? c.match[j - D] : 0.0;
    const int stop = -min(max(0, base_len - D), B);
    __syncwarp();
    int pos;
    const double pr = stut_walk<true, TRACE>(c, ins_prog, stop, j, k, lp0, B, pos);
    if (live) tcol[(HIPSTR_MAX_ARTIFACT_UNITS + k) * pitch] = __ldg(rep->art + HIPSTR_MAX_ARTIFACT_UNITS + k) + pr;
    if (TRACE && live) pcol[(HIPSTR_MAX_ARTIFACT_UNITS + k) * pitch] = pos;
  }
}

__host__ __device__ inline size_t stut_smem_bytes_hd(int n_max) {
  // val[5N] doubles, per warp match[N] + terms[SLOTS * 32] doubles, code[N] bytes, raw[2][2][N] bytes, 2 mbarriers,
  // job ids + task counter
  return ((size_t)HIPSTR_VAL_STRIDE * n_max + (size_t)STUT_WARPS * (n_max + STUT_TERM_SLOTS * 32)) * 8 + (size_t)n_max +
         4 * (size_t)n_max + 16 + 16;
}
size_t stutter_smem_bytes(int n_max) { return stut_smem_bytes_hd(n_max); }

// One CTA per job = (pooled read, up to 8 allele slots); its warps pull (slot, side) tasks -- left sides first, they
// are the long ones -- from a shared-memory counter, so no barrier separates the slots.
#ifndef STUT_MIN_CTAS
#define STUT_MIN_CTAS 7
#endif
template <bool TRACE>
__global__ void __launch_bounds__(STUT_THREADS, STUT_MIN_CTAS) k_stutter(const StutParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.n_max;   // multiple of 16
  double* s_val = reinterpret_cast<double*>(smem_raw);
  double* s_match = s_val + HIPSTR_VAL_STRIDE * N + warp * N;                         // this warp's
  double* s_terms = s_val + (HIPSTR_VAL_STRIDE + STUT_WARPS) * N + warp * STUT_TERM_SLOTS * 32;
  uint8_t* s_code = reinterpret_cast<uint8_t*>(s_val + (HIPSTR_VAL_STRIDE + STUT_WARPS) * N + STUT_WARPS * STUT_TERM_SLOTS * 32);
  uint8_t* s_raw = s_code + N;                       // [buffer][bases | quals][N]
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(s_raw);
  const unsigned bar_addr = raw_addr + 4 * N;        // two mbarriers
  volatile int* s_job = reinterpret_cast<volatile int*>(s_raw + 4 * N + 16);   // [0..1] job ids, [2] task counter
  int* s_task = const_cast<int*>(s_job) + 2;
  const unsigned val_addr = (unsigned)__cvta_generic_to_shared(s_val);
  const unsigned terms_addr = (unsigned)__cvta_generic_to_shared(s_terms + lane);

  if (tid == 0) { mbar_init(bar_addr, 1); mbar_init(bar_addr + 8, 1); }
  __syncthreads();
  // thread 0 pulls the next job and starts the TMA bulk copy of its read into landing zone `buf`
  auto fetch_and_stage = [&](int buf) {
    const int id = atomicAdd(P.job_counter, 1);
    s_job[buf] = id;
    if (id < P.n_jobs) {
      const DevPool* pp = P.pools + P.jobs[id].pool;
      const int off = pp->seq_off;
      const unsigned bytes = (unsigned)((pp->len + 15) / 16 * 16);
      const unsigned zone = raw_addr + buf * 2 * N;
      fence_async_smem();   // earlier generic reads of the zone are done
      mbar_expect_tx(bar_addr + 8 * buf, 2 * bytes);
      bulk_g2s(zone, P.bases + off, bytes, bar_addr + 8 * buf);
      bulk_g2s(zone + N, P.quals + off, bytes, bar_addr + 8 * buf);
    }
  };
  unsigned bar_phase[2] = {0, 0};
  int cur = 0;
  if (tid == 0) fetch_and_stage(0);
  __syncthreads();
  for (;;) {
    const int job_id = s_job[cur];
    if (job_id >= P.n_jobs) break;
    if (tid == 0) { fetch_and_stage(cur ^ 1); *s_task = 0; }   // every thread left the other zone before the barrier that ended the last job
    mbar_wait(bar_addr + 8 * cur, bar_phase[cur]);
    bar_phase[cur] ^= 1;
    const DevStutJob job = P.jobs[job_id];
    const DevPool pool = P.pools[job.pool];
    const int n = pool.len, seed = pool.seed;
    const int nL = seed, nR = n - seed - 1;
    // the emission table val[column][5] in SIDE order: left of the seed forwards, right of the seed reversed
    {
      const uint8_t* rawb = s_raw + cur * 2 * N;
      const uint8_t* rawq = rawb + N;
      for (int i = tid; i < n; i += STUT_THREADS) {
        if (i == seed) continue;
        const int g = i < seed ? i : nL + (n - 1 - i);
        const uint8_t q = rawq[i], x = rawb[i];
        const double ok = __ldg(P.qual_lut + 2 * q), bad = __ldg(P.qual_lut + 2 * q + 1);
        s_code[g] = x;
#pragma unroll
        for (int y = 0; y < 5; y++) s_val[g * HIPSTR_VAL_STRIDE + y] = (y == x) ? ok : bad;
      }
    }
    __syncthreads();
    const int pitch = hipstr_t_pitch(n);
    // alignment: one slab per pooled read; traces: one slab per job (a trace reads one haplotype)
    const int64_t t_off = P.job_t_off ? P.job_t_off[job_id] : P.pool_t_off[job.pool];
    double* tpool = P.stut + t_off;
    const int n_tasks = 2 * job.n_slots;
    for (;;) {
      int task = 0;
      if (lane == 0) task = atomicAdd(s_task, 1);
      task = __shfl_sync(FULL, task, 0);
      if (task >= n_tasks) break;
      const int side = task >= job.n_slots;              // every left side before the first right side
      const int s = side ? task - job.n_slots : task;
      const DevSlotReps sr = P.slot_reps[job.slot0 + s];
      const int n_side = side ? nR : nL;
      const int gbase = side ? nL : 0;
      double* tside = tpool + (size_t)(job.tslot0 + s) * HIPSTR_NUM_ARTIFACTS * pitch + gbase;
      int32_t* pside = TRACE ? P.stut_pos + t_off + (size_t)(job.tslot0 + s) * HIPSTR_NUM_ARTIFACTS * pitch + gbase : nullptr;
      const DevRep* rep = P.reps + (side ? sr.rep_rev : sr.rep_fwd);
      StutCtx c;
      c.progs = P.progs; c.rep = rep; c.int_logs = P.int_logs;
      c.diag = P.rep_tabs + __ldg(&rep->diag_off); c.ins_tab = P.rep_tabs + __ldg(&rep->ins_off);
      c.val = val_addr + gbase * HIPSTR_COL_BYTES; c.code = s_code + gbase; c.match = s_match;
      c.terms = terms_addr;
      c.B = __ldg(&rep->len); c.p = __ldg(&rep->period); c.n_side = n_side;
      const int B = c.B, p = c.p;
      const int n_del = min(HIPSTR_MAX_ARTIFACT_UNITS, B / p);
      // pass 1a: match_probs_ of every column q; the running sum after k * period terms is del_probs_[q][k-1], and
      // the first term of the deletion walk of column q - k * period is their difference (StutterAlignerClass.cpp:117-118)
      for (int q = lane; q < n_side; q += 32) {
        const unsigned col = c.val + q * HIPSTR_COL_BYTES;
        const int terms = min(q + 1, B);
        double acc = 0.0, snap[HIPSTR_MAX_ARTIFACT_UNITS];
#pragma unroll
        for (int k = 0; k < HIPSTR_MAX_ARTIFACT_UNITS; k++) snap[k] = 0.0;
        int t = 0;
#pragma unroll
        for (int k = 0; k < HIPSTR_MAX_ARTIFACT_UNITS; k++) {
          const int upto = min(terms, (k + 1) * p);
          for (; t < upto; t++) acc += lds_f64(col + __ldg(c.diag + t));
          snap[k] = acc;
        }
        for (; t < terms; t++) acc += lds_f64(col + __ldg(c.diag + t));
        s_match[q] = acc;
#pragma unroll
        for (int k = 0; k < HIPSTR_MAX_ARTIFACT_UNITS; k++)
          if (k < n_del && q - (k + 1) * p >= 0) tside[(HIPSTR_MAX_ARTIFACT_UNITS - 1 - k) * pitch + q - (k + 1) * p] = acc - snap[k];
      }
      // pass 1b: the columns whose deletion partner lies beyond the read end sum their first term directly:
      // read base j-t against allele base B-1-(t-D), starting from the prior (StutterAlignerClass.cpp:119-121)
      {
        int task_base = 0;
        for (int k = 1; k <= n_del; k++) {
          const int D = -k * p;
          const int cnt = min(-D, n_side);                 // columns n_side - cnt .. n_side - 1
          for (int u = lane - (task_base & 31); u < cnt; u += 32) {
            if (u < 0) continue;
            const int j = n_side - cnt + u;
            const int base_len = min(B + D, j + 1);
            const unsigned cold = c.val + (j - D) * HIPSTR_COL_BYTES;
            const int32_t* dg = c.diag - D;
            double v = -__ldg(c.int_logs + (B + D + 1));
#pragma unroll 4
            for (int t = 0; t < base_len; t++) v += lds_f64(cold + __ldg(dg + t));
            tside[(HIPSTR_MAX_ARTIFACT_UNITS - k) * pitch + j] = v;
          }
          task_base += cnt;
        }
      }
      __syncwarp();
      // pass 2: the 12 walks + the no-artifact entry of every column
      for (int j0 = 0; j0 < n_side; j0 += 32) {
        const int jq = j0 + lane;
        const bool live = jq < n_side;
        const int j = live ? jq : n_side - 1;   // idle lanes shadow the side's last column (the warp stays converged)
        __syncwarp();
        stut_column<TRACE>(c, j, live, tside + j, TRACE ? pside + j : nullptr, pitch);
      }
      __syncwarp();   // the next task overwrites this warp's tables
    }
    __syncthreads();
    cur ^= 1;
  }
}

template <bool TRACE>
static cudaError_t launch_stutter_t(const StutParams& p, cudaStream_t stream) {
  const size_t smem = stut_smem_bytes_hd(p.n_max);
  cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void*>(&k_stutter<TRACE>));
  if (e != cudaSuccess) return e;
  const int sms = sm_count_of_current_device();
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stutter<TRACE>, STUT_THREADS, smem);
  if (e != cudaSuccess) return e;
  // persistent grid: exactly as many CTAs as can be resident (a multiple of the SM count)
  int grid = sms * (per_sm > 0 ? per_sm : 1);
  if (grid > p.n_jobs) grid = p.n_jobs;
  k_stutter<TRACE><<<grid, STUT_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_stutter(const StutParams& p, cudaStream_t stream) {
  if (p.n_jobs <= 0) return cudaSuccess;
  return p.stut_pos ? launch_stutter_t<true>(p, stream) : launch_stutter_t<false>(p, stream);
}

}  // namespace hipstr
