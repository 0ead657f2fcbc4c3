/*
 * kernels.cu -- hand-written sm_100a kernels of the HipSTR hot path.
 *
 * K1  k_align<C>      one warp per (pooled read, run of haplotypes): the seeded two-sided
 *                     read-vs-haplotype HMM of HapAligner::process_read
 *                     (SeqAlignment/HapAligner.cpp:573-709 -> align_seq_to_hap :26-161,
 *                     compute_aln_logprob :163-231, StutterAlignerClass.cpp:12-162).
 * K2  k_scatter       pool -> read scatter + mate merge (seq_stutter_genotyper.cpp:530-564).
 * K3  k_posteriors    genotype posteriors (genotyper.cpp:20-97).
 *
 * K1 in one paragraph.  The read is split at its seed base into a left part aligned to the
 * forward haplotype and a right part aligned (reversed) to the reversed haplotype.  Both DPs run
 * AT THE SAME TIME in one warp: the lanes are partitioned between the two sides in proportion to
 * their column counts, every lane owns C adjacent read columns of its side and keeps the previous
 * DP row of those columns (match + deletion state; the insertion state only flows along a row) in
 * registers.  Flank rows advance as an anti-diagonal wavefront: at step t lane k computes
 * haplotype row t-k for its columns and hands the right-most cell to lane k+1 with three warp
 * shuffles.  A repeat ("stutter") block is a single super-row: the row above it is parked in
 * shared memory, every lane evaluates the 13 PCR-artifact sizes for one read column at a time
 * (the per-artifact sums over artifact positions of StutterAlignerClass), and the row is pulled
 * back into registers.  Only the last column of every row is kept (shared memory); the final
 * log-likelihood combines them over all seed placements.  All state is FP64 and every operation
 * is done in the reference's order, so flank cells are bit-identical to the CPU; the approximate
 * log-sum-exp uses the float replicas in fastapprox.cuh.  The double sums inside a log-sum-exp
 * add float values spanning < 2^11 in magnitude, which is exact in binary64 in any order, so
 * warp-parallel reductions of those sums do not change a bit either.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <stdlib.h>

#include "../../include/hipstr_b200.h"
#include "fastapprox.cuh"
#include "kcommon.cuh"
#include "kernels.h"
#include "layout.h"

namespace hipstr {

#define LOG_INS_TO_INS (-1.0)                 /* AlignmentModel.h:7 */
#define LOG_INS_TO_MATCH (-0.4586751453870818910216436) /* AlignmentModel.h:8 */
#define LOG_DEL_TO_DEL (-1.0)                 /* AlignmentModel.h:9 */
#define LOG_DEL_TO_MATCH (-0.4586751453870818910216436) /* AlignmentModel.h:10 */

__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(FULL, v, 1); }

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t align_smem_doubles(int n_max, int l_max) {
  // run[N] rowout[N] rowbuf[N] + 2 x raw[2][round16(N)] bytes (double-buffered bulk-copy landing zones) + two mbarriers
  (void)l_max;
  const size_t n16 = ((size_t)n_max + 15) / 16 * 16;
  const size_t d = 3 * (size_t)n16 + 4 * n16 / 8 + 2;
  return (d + 1) / 2 * 2;   // keep every warp's slab 16-byte aligned
}
size_t align_smem_bytes(int n_max, int l_max) { return align_smem_doubles(n_max, l_max) * 8 * HIPSTR_WARPS_PER_CTA; }

// TRACE = the forward pass of K5 (HapAligner::trace_optimal_aln, HapAligner.cpp:711-722 -> process_read(retrace_aln)):
// one job per trace = (pooled read, ONE haplotype).  The reference keeps the three full matrices and re-derives the best
// predecessor of every cell it visits walking back; those choices only depend on values at hand when the cell is
// computed, so they are taken HERE (same tolerance and side-dependent preferences, kcommon.cuh pick2 / pick3) and stored
// as one byte per flank cell, plus the best artifact size / position of every repeat-block column; k_trace_walk
// (trace.cu) then follows the bytes.
template <int C, int MINB, bool TRACE>
__global__ void __launch_bounds__(32 * HIPSTR_WARPS_PER_CTA, MINB) k_align(const AlignParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int N = P.n_max, L = P.l_max;
  // last-column values of both sides live in a per-warp slab of global memory (L1/L2 resident):
  // they are written once per row and read once per haplotype, not worth shared memory
  double* s_last = P.last_scratch + ((size_t)blockIdx.x * HIPSTR_WARPS_PER_CTA + wib) * 2 * (size_t)L;
  // Persistent warps: every warp pulls (pooled read, haplotype range) jobs from a global counter.
  const int N16 = (N + 15) / 16 * 16;
  double* wbase0 = reinterpret_cast<double*>(smem_raw) + (size_t)wib * align_smem_doubles(N, L);
  uint8_t* s_raw0 = reinterpret_cast<uint8_t*>(wbase0 + 3 * (size_t)N16);
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(s_raw0);
  const unsigned bar_addr = raw_addr + 4 * N16;   // two 8-byte mbarriers after the two zones
  if (lane == 0) { mbar_init(bar_addr, 1); mbar_init(bar_addr + 8, 1); }
  __syncwarp();
  unsigned bar_phase[2] = {0, 0};
  // Double-buffered TMA staging: the bulk copy of the NEXT job's read (packed bases + qualities) is
  // issued before the current job is processed, so its latency never shows.
  auto fetch_and_stage = [&](int buf) {
    int id = 0;
    if (lane == 0) {
      id = atomicAdd(P.job_counter, 1);
      if (id < P.n_jobs) {
        const DevPool* pp = P.pools + P.jobs[id].pool;
        const int len = pp->len, off = pp->seq_off;
        const unsigned bytes = (unsigned)((len + 15) / 16 * 16);
        const unsigned zone = raw_addr + buf * 2 * N16;
        if (pp->seed < 0 || (int)bytes > N16) {
          // a read without a seed is never aligned (and may be longer than this launch's zones):
          // complete the barrier phase without a copy
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar_addr + 8 * buf) : "memory");
        } else {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the zone are done
          mbar_expect_tx(bar_addr + 8 * buf, 2 * bytes);
          bulk_g2s(zone, P.bases + off, bytes, bar_addr + 8 * buf);
          bulk_g2s(zone + N16, P.quals + off, bytes, bar_addr + 8 * buf);
        }
      }
    }
    return __shfl_sync(FULL, id, 0);
  };
  int cur_buf = 0;
  int job_id = fetch_and_stage(0);
  for (;;) {
  if (job_id >= P.n_jobs) break;
  __syncwarp();   // every lane is done with the zone the next copy will overwrite
  const int next_job_id = fetch_and_stage(cur_buf ^ 1);
  const uint8_t* s_rawb = s_raw0 + cur_buf * 2 * N16;
  const uint8_t* s_rawq = s_rawb + N16;
  mbar_wait(bar_addr + 8 * cur_buf, bar_phase[cur_buf]);
  bar_phase[cur_buf] ^= 1;
  const int this_buf = cur_buf;
  (void)this_buf;
  cur_buf ^= 1;
  const int my_job = job_id;
  job_id = next_job_id;
  const DevJob job = P.jobs[my_job];
  const DevPool pool = P.pools[job.pool];
  double* out = P.ll_out + pool.out_off;

  if (pool.seed < 0) {   // HapAligner.cpp:333-337
    for (int h = lane; h < pool.n_haps; h += 32) {
      out[h] = 0.0;
      if (P.pos_out) P.pos_out[pool.out_off + h] = -1;
    }
    continue;
  }

  double* wbase = reinterpret_cast<double*>(smem_raw) + (size_t)wib * align_smem_doubles(N, L);
  double* s_run = wbase;               // running sums of log_correct (row 0 of every haplotype starts from them)
  double* s_rowbuf = s_run + N16;      // the row above a repeat block
  double* s_rowout = s_rowbuf + N16;   // the repeat block's output row / the parked deletion row

  const int n = pool.len, seed = pool.seed;
  const int nL = seed, nR = n - seed - 1;
  // Columns are in SIDE order: columns 0..nL-1 of the left side are read bases 0..seed-1 (aligned to the forward
  // haplotype), columns 0..nR-1 of the right side are read bases n-1..seed+1 (reversed, aligned to the reversed
  // haplotype), HapAligner.cpp:579-585,606-609.  The packed bases and qualities arrived by TMA bulk copy in the
  // landing zone (waited for above).
  const uint8_t seed_code = s_rawb[seed];
  const uint8_t seed_q = s_rawq[seed];
  const double seed_ok = __ldg(P.qual_lut + 2 * seed_q), seed_bad = __ldg(P.qual_lut + 2 * seed_q + 1);
  // running sums of log_correct from each read end towards the seed (row 0 of either matrix,
  // HapAligner.cpp:33-42); strictly sequential adds, one lane per side
  double edge = 0.0;
  if (lane < 2) {
    const int cnt = lane ? nR : nL;
    const int g0 = lane ? nL : 0;
    double acc = 0.0;
#pragma unroll 4
    for (int j = 0; j < cnt; j++) {
      s_run[g0 + j] = acc;
      acc += __ldg(P.qual_lut + 2 * s_rawq[lane ? n - 1 - j : j]);
    }
    edge = acc;
  }
  const double edgeL = __shfl_sync(FULL, edge, 0), edgeR = __shfl_sync(FULL, edge, 1);
  __syncwarp();

  // lane -> (side, first column)
  const int nlL = (nL + C - 1) / C, nlR = (nR + C - 1) / C;
  const int side = lane < nlL ? 0 : 1;
  const int k = side ? lane - nlL : lane;
  const bool lane_on = side == 0 || k < nlR;
  const int ncol = side ? nR : nL;
  const int gbase = side ? nL : 0;
  const int j0 = k * C;

  // per-column constants stay in registers for the whole job: base code, log P(correct), log P(error)
  double lc[C], lw[C], Mp[C], Dp[C];
  uint8_t bs[C];
#pragma unroll
  for (int cc = 0; cc < C; cc++) {
    const int j = j0 + cc;
    const bool ok = lane_on && j < ncol;
    const int i = ok ? (side ? n - 1 - j : j) : 0;
    const uint8_t q = s_rawq[i];
    bs[cc] = s_rawb[i];
    lc[cc] = __ldg(P.qual_lut + 2 * q);
    lw[cc] = __ldg(P.qual_lut + 2 * q + 1);
  }
  const int last_cc = (lane_on && ncol - 1 >= j0 && ncol - 1 < j0 + C) ? ncol - 1 - j0 : -1;
  __syncwarp();
  const int t_pitch = hipstr_t_pitch(n);
  const int64_t t_off = TRACE ? P.job_t_off[my_job] : P.pool_t_off[job.pool];
  const double* t_pool = P.stut + t_off;
  const int tr = job.pad;   // TRACE: the trace this job computes

  // The stutter tables come from HBM (K1a wrote them a moment ago): ask L2 for the tables of a haplotype one
  // haplotype ahead of their use, 128 bytes per lane and request, so that the 13 loads per column find them on chip.
  auto prefetch_tables = [&](int h) {
    if (TRACE || h >= job.h1 || (P.hap_mask && !P.hap_mask[(pool.hap_rec0 >> 1) + h])) return;
    const DevHapSide hp = P.hapsides[pool.hap_rec0 + 2 * h];
    for (int b = 0; b < hp.n_blocks; b++) {
      const DevBlock blk = P.blocks[hp.blk_off + b];
      if (blk.rep < 0) continue;
      const char* slab = reinterpret_cast<const char*>(t_pool + (size_t)blk.tslot * HIPSTR_NUM_ARTIFACTS * t_pitch);
      const int bytes = HIPSTR_NUM_ARTIFACTS * t_pitch * 8;
      for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(slab + o));
    }
  };
  prefetch_tables(job.h0);

  int cached_class = -1;   // seg1_class of the rows before the first repeat block that this lane's
                           // side currently holds in s_rowbuf / s_last (from an earlier haplotype)
  for (int h = job.h0; h < job.h1; h++) {
    const int hap_index = (pool.hap_rec0 >> 1) + h;
    prefetch_tables(h + 1);
    if (P.hap_mask && !P.hap_mask[hap_index]) continue;
    const DevHapSide hsF = P.hapsides[pool.hap_rec0 + 2 * h];
    const DevHapSide hsR = P.hapsides[pool.hap_rec0 + 2 * h + 1];
    const DevHapSide& hs = side ? hsR : hsF;
    const uint8_t* seq = P.hapbytes + hs.seq_off;
    const uint8_t* rows = P.hapbytes + hs.row_off;
    const int nb = hsF.n_blocks;
    const int hlen = hsF.len;
    unsigned char* dec_side = TRACE ? P.dec + P.dec_off[tr] + (side ? (size_t)hlen * nL : 0) : nullptr;
    int32_t* art_base = TRACE ? P.art + P.art_off[tr] : nullptr;   // sizes [2 sides][blocks][n_side], then positions

    // Rows before the first repeat block depend only on the read and on seg1_class: when the
    // previous haplotype of this job had the same class they are still in shared memory (the row
    // above the repeat block in s_rowbuf, the last-column values in s_last) -- the same reuse the
    // reference gets from walking haplotypes in Gray-code order (HapAligner.cpp:54-60).
    const bool reuse = hs.seg1_class >= 0 && hs.seg1_class == cached_class;
    const int first_rep = hs.first_rep;
    cached_class = hs.seg1_class;
    // row 0 (HapAligner.cpp:33-42)
    if (!reuse) {
      const uint8_t fc = __ldg(seq);
#pragma unroll
      for (int cc = 0; cc < C; cc++) {
        const int j = j0 + cc;
        Mp[cc] = (bs[cc] == fc ? lc[cc] : lw[cc]) + ((lane_on && j < ncol) ? s_run[gbase + j] : 0.0);
        Dp[cc] = IMPOSSIBLE;
        if (cc == last_cc) s_last[side * L] = Mp[cc];
      }
    }

    for (int b = 0; b < nb; b++) {
      const DevBlock blkF = P.blocks[hsF.blk_off + b];
      const DevBlock blkR = P.blocks[hsR.blk_off + b];
      const DevBlock& blk = side ? blkR : blkF;

      // ---------------- flank block: anti-diagonal wavefront (HapAligner.cpp:110-157) ----------
      {
        const int r0 = blk.row_start + (b == 0 ? 1 : 0);
        const int nrows = blk.row_start + blk.len - r0;
        const bool flank_on = lane_on && blk.rep < 0 && nrows > 0 && !(reuse && b < first_rep);
        const int steps = __reduce_max_sync(FULL, flank_on ? nrows + k : 0);
        if (steps > 0) {
          double Mlp = shfl_up_d(Mp[C - 1]), Dlp = shfl_up_d(Dp[C - 1]);
          double pubM = 0.0, pubI = 0.0, pubD = 0.0;
          for (int t = 0; t < steps; t++) {
            const double rI = shfl_up_d(pubI), rM = shfl_up_d(pubM), rD = shfl_up_d(pubD);
            const int r = t - k;
            const bool live = flank_on && r >= 0 && r < nrows;
            const int row = r0 + r;
            const uint8_t info = live ? __ldg(rows + row) : 0;
            // the first row after a repeat block is a different recurrence (HapAligner.cpp:129-139); one row per side and
            // haplotype is such a row, so the cells carry its selects only in the steps where some lane stands on one
            const bool some_after = __any_sync(FULL, (info & HIPSTR_ROW_AFTER_REPEAT) != 0);
            if (live) {
              const uint8_t hc = __ldg(seq + row);
              const int hp = info & 15;
              const bool after = (info & HIPSTR_ROW_AFTER_REPEAT) != 0;
              const double m2m = __ldg(P.trans + hp), m2i = __ldg(P.trans + 16 + hp), m2d = __ldg(P.trans + 32 + hp);
              double Ileft = rI, Mdiag = Mlp, Ddiag = Dlp;
              auto cells = [&](auto with_after) {
#pragma unroll
                for (int cc = 0; cc < C; cc++) {
                  const double e = bs[cc] == hc ? lc[cc] : lw[cc];
                  const double Mup = Mp[cc], Dup = Dp[cc];
                  const bool col0 = (cc == 0) && (k == 0);
                  double Mn = e + dmax(Ileft + m2i, dmax(Mdiag + m2m, Ddiag + m2d));
                  double In = lc[cc] + dmax(Mdiag + LOG_INS_TO_MATCH, Ileft + LOG_INS_TO_INS);
                  double Dn = dmax(Mup + LOG_DEL_TO_MATCH, Dup + LOG_DEL_TO_DEL);
                  if (col0) { Mn = e; In = lc[cc]; }
                  if (decltype(with_after)::value && after) {
                    Mn = col0 ? e : e + Mdiag;
                    In = IMPOSSIBLE;
                    Dn = IMPOSSIBLE;
                  }
                  if (TRACE && j0 + cc < ncol) {   // what retrace would choose standing on this cell, from the very values it would read
                    const bool rev = side != 0;
                    const int choice = col0 ? (pick2(rev, Dup + LOG_DEL_TO_DEL, Mup + LOG_DEL_TO_MATCH) << 2)
                                            : (pick3(rev, Ileft + m2i, Ddiag + m2d, Mdiag + m2m) |
                                               (pick2(rev, Dup + LOG_DEL_TO_DEL, Mup + LOG_DEL_TO_MATCH) << 2) |
                                               (pick2(rev, Ileft + LOG_INS_TO_INS, Mdiag + LOG_INS_TO_MATCH) << 3));
                    dec_side[(size_t)row * ncol + j0 + cc] = (unsigned char)choice;
                  }
                  Mdiag = Mup; Ddiag = Dup; Ileft = In;
                  Mp[cc] = Mn; Dp[cc] = Dn;
                  if (cc == last_cc) s_last[side * L + row] = Mn;
                }
              };
              if (some_after) cells(std::true_type()); else cells(std::false_type());
              Mlp = rM; Dlp = rD;
              pubM = Mp[C - 1]; pubI = Ileft; pubD = Dp[C - 1];
            }
          }
        }
      }

      // ---------------- repeat block: one super-row (HapAligner.cpp:62-109) --------------------
      if (blkF.rep >= 0 || blkR.rep >= 0) {
        if (lane_on && !(reuse && b == first_rep)) {   // park the row above the block (a side that is in a
                                                        // flank block parks M and D); a reused row is already there
#pragma unroll
          for (int cc = 0; cc < C; cc++)
            if (j0 + cc < ncol) {
              s_rowbuf[gbase + j0 + cc] = Mp[cc];
              if (blk.rep < 0) s_rowout[gbase + j0 + cc] = Dp[cc];
            }
        }
        __syncwarp();
        // every column of a side that is in a repeat block: fold the stutter tables of K1a with the row above
        // (HapAligner.cpp:84-100): probs[D] = T[D][column] + M[row above][j - base_len], then the 13-term log-sum-exp
        for (int g = lane; g < n - 1; g += 32) {
          const int gs = g >= nL;
          const DevBlock& gb = gs ? blkR : blkF;
          if (gb.rep < 0) continue;
          const DevRep* rep = P.reps + gb.rep;
          const int B = __ldg(&rep->len), p = __ldg(&rep->period);
          const int j = g - (gs ? nL : 0);
          int tslot = gb.tslot;
          if (TRACE) {   // the trace's slab holds one table per repeat block of the haplotype, in forward block order
            const int fb = gs ? nb - 1 - b : b;
            tslot = 0;
            for (int x = 0; x < fb; x++) tslot += P.blocks[hsF.blk_off + x].rep >= 0;
          }
          const double* tcol = t_pool + (size_t)tslot * HIPSTR_NUM_ARTIFACTS * t_pitch + g;
          const double* prev = s_rowbuf + (gs ? nL : 0);
          // all 13 table loads first, with nothing between them that waits for one: they are in flight together.  The
          // entries of impossible sizes are allocated but never written (K1a skips them); they are loaded and dropped.
          double probs[HIPSTR_NUM_ARTIFACTS];
#pragma unroll
          for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) probs[a] = __ldg(tcol + a * t_pitch);
#pragma unroll
          for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) {
            const int BD = B + (a - HIPSTR_MAX_ARTIFACT_UNITS) * p;
            const int from = j - min(BD, j + 1);                   // column of the row above the block this size continues from
            const double above = prev[max(from, 0)];
            probs[a] = select_if_nonneg(BD, probs[a] + select_if_nonneg(from, above, 0.0), IMPOSSIBLE);
          }
          double mx = probs[0];
#pragma unroll
          for (int a = 1; a < HIPSTR_NUM_ARTIFACTS; a++) mx = dmax(mx, probs[a]);
          double total = 0.0;
#pragma unroll
          for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) total += lse_term(probs[a], mx);
          s_rowout[g] = lse_finish(mx, total);
          if (TRACE) {   // HapAligner.cpp:79-97: the first strictly best artifact size and its position
            double best = IMPOSSIBLE;
            int best_a = -1;
#pragma unroll
            for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++)
              if (probs[a] > best) { best = probs[a]; best_a = a; }
            const int n_side_g = gs ? nR : nL;
            int32_t* sizes = art_base + (gs ? nb * nL : 0) + b * n_side_g;
            int32_t* poss = sizes + nb * (nL + nR);
            sizes[j] = best_a < 0 ? -10000 : (best_a - HIPSTR_MAX_ARTIFACT_UNITS) * p;
            poss[j] = best_a < 0 ? 0 : __ldg(P.stut_pos + t_off + (size_t)tslot * HIPSTR_NUM_ARTIFACTS * t_pitch + g + (size_t)best_a * t_pitch);
          }
        }
        __syncwarp();
        // back to registers
#pragma unroll
        for (int cc = 0; cc < C; cc++) {
          const int j = j0 + cc;
          const bool ok = lane_on && j < ncol;
          const int g = ok ? gbase + j : 0;
          if (blk.rep >= 0) {
            Mp[cc] = ok ? s_rowout[g] : 0.0;
            Dp[cc] = IMPOSSIBLE;
            if (cc == last_cc) s_last[side * L + blk.row_start + blk.len - 1] = Mp[cc];
          } else {   // this side is not in a repeat block at this block index: state was parked
            Mp[cc] = ok ? s_rowbuf[g] : 0.0;
            Dp[cc] = ok ? s_rowout[g] : 0.0;
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();

    // ---------------- combine over seed placements (HapAligner.cpp:163-231) -------------------
    {
      const uint8_t* fseq = P.hapbytes + hsF.seq_off;
      const uint8_t* frow = P.hapbytes + hsF.row_off;
      const double prior = -__ldg(P.int_logs + hsF.n_seed_pos);
      const double* lastL = s_last;
      const double* lastR = s_last + L;
      double vmax = -1.0e300;
      int vrank = 0x7fffffff;
      // rank 0: seed on haplotype base 0; rank 1: seed on the last base; rank i+1: interior base i
      for (int it = lane; it < hlen; it += 32) {
        double v;
        int rank;
        if (it == 0) {
          v = prior + (seed_code == __ldg(fseq) ? seed_ok : seed_bad) + edgeL + lastR[hlen - 2];
          rank = 0;
        } else if (it == hlen - 1) {
          v = prior + (seed_code == __ldg(fseq + hlen - 1) ? seed_ok : seed_bad) + edgeR + lastL[hlen - 2];
          rank = 1;
        } else {
          if (__ldg(frow + it) & HIPSTR_ROW_REPEAT) continue;
          v = prior + (seed_code == __ldg(fseq + it) ? seed_ok : seed_bad) + lastL[it - 1] + lastR[hlen - it - 2];
          rank = it + 1;
        }
        if (v > vmax || (v == vmax && rank < vrank)) { vmax = v; vrank = rank; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, vmax, o);
        const int orank = __shfl_xor_sync(FULL, vrank, o);
        if (ov > vmax || (ov == vmax && orank < vrank)) { vmax = ov; vrank = orank; }
      }
      double total = 0.0;
      for (int it = lane; it < hlen; it += 32) {
        double v;
        if (it == 0)
          v = prior + (seed_code == __ldg(fseq) ? seed_ok : seed_bad) + edgeL + lastR[hlen - 2];
        else if (it == hlen - 1)
          v = prior + (seed_code == __ldg(fseq + hlen - 1) ? seed_ok : seed_bad) + edgeR + lastL[hlen - 2];
        else {
          if (__ldg(frow + it) & HIPSTR_ROW_REPEAT) continue;
          v = prior + (seed_code == __ldg(fseq + it) ? seed_ok : seed_bad) + lastL[it - 1] + lastR[hlen - it - 2];
        }
        total += lse_term(v, vmax);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
      if (lane == 0) {
        const int seed_pos = vrank == 0 ? 0 : (vrank == 1 ? hlen - 1 : vrank - 1);
        if (TRACE) P.trace_seed_pos[tr] = seed_pos;
        else {
          out[h] = lse_finish(vmax, total);
          if (P.pos_out) P.pos_out[pool.out_off + h] = seed_pos;
        }
      }
      if (P.debug_out && my_job == 0 && h == job.h1 - 1)
        for (int i = lane; i < 2 * L; i += 32) P.debug_out[i] = (i % L) < hlen ? s_last[i] : 0.0;
    }
    __syncwarp();
  }
  __syncwarp();
  }   // next job
}

template <int C, int MINB, bool TRACE = false>
static cudaError_t launch_align_c(const AlignParams& p, int max_ctas, cudaStream_t stream, int* grid_out) {
  const size_t smem = align_smem_bytes(p.n_max, p.l_max);
  cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void*>(&k_align<C, MINB, TRACE>));
  if (e != cudaSuccess) return e;
  const int sms = sm_count_of_current_device();
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align<C, MINB, TRACE>, 32 * HIPSTR_WARPS_PER_CTA, smem);
  if (e != cudaSuccess) return e;
  // persistent grid: exactly as many CTAs as can be resident (a multiple of the SM count)
  const int jobs_ctas = (p.n_jobs + HIPSTR_WARPS_PER_CTA - 1) / HIPSTR_WARPS_PER_CTA;
  int grid = sms * (per_sm > 0 ? per_sm : 1);
  if (grid > jobs_ctas) grid = jobs_ctas;
  if (grid > max_ctas) grid = max_ctas;
  if (grid_out) *grid_out = grid;
  k_align<C, MINB, TRACE><<<grid, 32 * HIPSTR_WARPS_PER_CTA, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_align(int variant, const AlignParams& p, int max_ctas, cudaStream_t stream, int* grid_out) {
  if (p.n_jobs <= 0) return cudaSuccess;
  // tuning switch: HIPSTR_ALIGN_WARPS=20 caps the short-read variants at 96 registers (20 one-warp CTAs per SM)
  static const int dense = [] { const char* e = getenv("HIPSTR_ALIGN_WARPS"); return e && atoi(e) >= 20; }();
  switch (variant) {
    case 0: return dense ? launch_align_c<2, 20>(p, max_ctas, stream, grid_out) : launch_align_c<2, 16>(p, max_ctas, stream, grid_out);
    case 1: return dense ? launch_align_c<3, 20>(p, max_ctas, stream, grid_out) : launch_align_c<3, 16>(p, max_ctas, stream, grid_out);
    case 2: return dense ? launch_align_c<4, 20>(p, max_ctas, stream, grid_out) : launch_align_c<4, 16>(p, max_ctas, stream, grid_out);
    case 3: return dense ? launch_align_c<5, 20>(p, max_ctas, stream, grid_out) : launch_align_c<5, 16>(p, max_ctas, stream, grid_out);
    case 4: return launch_align_c<6, 12>(p, max_ctas, stream, grid_out);
    case 5: return launch_align_c<8, 12>(p, max_ctas, stream, grid_out);
    case 6: return launch_align_c<12, 8>(p, max_ctas, stream, grid_out);
    case 7: return launch_align_c<16, 8>(p, max_ctas, stream, grid_out);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_trace_forward(int variant, const AlignParams& p, int max_ctas, cudaStream_t stream) {
  if (p.n_jobs <= 0) return cudaSuccess;
  switch (variant) {
    case 0: return launch_align_c<2, 16, true>(p, max_ctas, stream, nullptr);
    case 1: return launch_align_c<3, 16, true>(p, max_ctas, stream, nullptr);
    case 2: return launch_align_c<4, 16, true>(p, max_ctas, stream, nullptr);
    case 3: return launch_align_c<5, 16, true>(p, max_ctas, stream, nullptr);
    case 4: return launch_align_c<6, 12, true>(p, max_ctas, stream, nullptr);
    case 5: return launch_align_c<8, 12, true>(p, max_ctas, stream, nullptr);
    case 6: return launch_align_c<12, 8, true>(p, max_ctas, stream, nullptr);
    case 7: return launch_align_c<16, 8, true>(p, max_ctas, stream, nullptr);
  }
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// K2: pool -> read scatter + mate merge.  One thread per (chain head read, haplotype): a chain is
// a read followed by its second mates, processed in order exactly like the reference's two loops.
// ------------------------------------------------------------------------------------------
__global__ void k_scatter(const ScatterParams P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)P.n_reads * P.n_haps) return;
  const int r = (int)(idx / P.n_haps), h = (int)(idx % P.n_haps);
  if (P.second_mate[r] && r > 0) return;   // handled by its chain head
  const bool hap_on = !P.realign_hap || P.realign_hap[h];
  for (int m = r; m < P.n_reads && (m == r || P.second_mate[m]); m++) {
    const bool copy = !P.copy_read || P.copy_read[m];
    if (!copy) continue;
    const int pi = P.pool_index[m];
    if (h == 0 && P.read_seed) P.read_seed[m] = P.pool_seed[pi];
    if (!hap_on) continue;
    double v = P.pool_ll[(size_t)pi * P.n_haps + h];
    if (m > r) {
      v += P.read_ll[(size_t)(m - 1) * P.n_haps + h];
      P.read_ll[(size_t)(m - 1) * P.n_haps + h] = v;
    }
    P.read_ll[(size_t)m * P.n_haps + h] = v;
  }
}

cudaError_t launch_scatter(const ScatterParams& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_reads * p.n_haps;
  if (total <= 0) return cudaSuccess;
  k_scatter<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

__global__ void k_scatter_batch(const ScatterBatchParams P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.n_elems) return;
  int lo = 0, hi = P.n_loci - 1;   // last locus with elem_off <= idx
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.loci[mid].elem_off <= idx) lo = mid; else hi = mid - 1;
  }
  const ScatterLocus L = P.loci[lo];
  const int64_t e = idx - L.elem_off;
  const int r = (int)(e / L.n_haps), h = (int)(e % L.n_haps);
  const uint8_t* second = P.second_mate + L.read0;
  if (second[r] && r > 0) return;
  const bool hap_on = !P.hap_mask || P.hap_mask[L.hap0 + h];
  const int32_t* pidx = P.pool_index + L.read0;
  double* ll = P.read_ll + L.elem_off;
  const double* pll = P.pool_ll + L.pool_ll_off;
  for (int m = r; m < L.n_reads && (m == r || second[m]); m++) {
    const int pi = pidx[m];
    bool copy = !P.copy_read || P.copy_read[L.read0 + m];
    if (!copy) continue;
    if (h == 0 && P.read_seed) P.read_seed[L.read0 + m] = P.pool_seed[L.pool0 + pi];
    if (!hap_on) continue;
    double v = pll[(size_t)pi * L.n_haps + h];
    if (m > r) {
      v += ll[(size_t)(m - 1) * L.n_haps + h];
      ll[(size_t)(m - 1) * L.n_haps + h] = v;
    }
    ll[(size_t)m * L.n_haps + h] = v;
  }
}

cudaError_t launch_scatter_batch(const ScatterBatchParams& p, cudaStream_t stream) {
  if (p.n_elems <= 0) return cudaSuccess;
  k_scatter_batch<<<(unsigned)((p.n_elems + 255) / 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K3: genotype posteriors.  One CTA per (locus, sample); thread t owns diplotypes t, t+T, ...
// and folds the sample's reads into them in read order (genotyper.cpp:59-64), then the CTA
// normalises with an exact log-sum-exp (genotyper.cpp:66-71) and picks the first maximum
// (genotyper.cpp:82-97).
// ------------------------------------------------------------------------------------------
#define POST_THREADS 128

__device__ __forceinline__ double block_reduce_max(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmax(v, __shfl_xor_sync(FULL, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < POST_THREADS / 32; w++) r = dmax(r, scratch[w]);
  return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < POST_THREADS / 32; w++) r += scratch[w];
  return r;
}

__global__ void __launch_bounds__(POST_THREADS) k_posteriors(const PostParams P) {
  __shared__ double scratch[POST_THREADS / 32];
  __shared__ int best_idx_s[POST_THREADS / 32];
  const PostSample S = P.samples[blockIdx.x];
  const int H = S.n_haps;
  const int64_t HH = (int64_t)H * H;
  double* post = P.post_out + S.post_off;
  double homoz, hetz;
  if (S.haploid) { homoz = -P.int_logs[H]; hetz = -1.7976931348623157e308 / 2; }
  else { homoz = P.int_logs[2] - P.int_logs[H] - P.int_logs[H + 1]; hetz = -P.int_logs[H] - P.int_logs[H + 1]; }
  const double* ll0 = P.read_ll + S.ll_off;
  double mx = -1.7976931348623157e308;
  for (int64_t d = threadIdx.x; d < HH; d += POST_THREADS) {
    const int a = (int)(d / H), b = (int)(d % H);
    double acc = a == b ? homoz : hetz;
    for (int r = S.read0; r < S.read1; r++) {
      const double* row = ll0 + (size_t)(r - S.locus_read0) * H;
      const double x = P.log_one_half + P.log_p1[r] + row[a];
      const double y = P.log_one_half + P.log_p2[r] + row[b];
      acc += P.read_weight[r] * lse2(x, y);
    }
    post[d] = acc;
    mx = dmax(mx, acc);
  }
  mx = block_reduce_max(mx, scratch);
  double sum = 0.0;
  for (int64_t d = threadIdx.x; d < HH; d += POST_THREADS) sum += exp(post[d] - mx);
  sum = block_reduce_sum(sum, scratch);
  const double sll = mx + log(sum);
  double best = -1.7976931348623157e308;
  int64_t best_d = HH;
  for (int64_t d = threadIdx.x; d < HH; d += POST_THREADS) {
    const double v = post[d] - sll;
    post[d] = v;
    if (v > best) { best = v; best_d = d; }   // ascending d per thread: first maximum kept
  }
  if (threadIdx.x == 0) P.sample_ll_out[blockIdx.x] = sll;
  if (P.best_out) {
    // first maximum over the CTA: larger value wins, ties go to the smaller index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(FULL, best, o);
      const long long od = __shfl_xor_sync(FULL, (long long)best_d, o);
      if (ov > best || (ov == best && od < best_d)) { best = ov; best_d = od; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = best; best_idx_s[threadIdx.x >> 5] = (int)best_d; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double bv = scratch[0];
      int bd = best_idx_s[0];
      for (int w = 1; w < POST_THREADS / 32; w++)
        if (scratch[w] > bv || (scratch[w] == bv && best_idx_s[w] < bd)) { bv = scratch[w]; bd = best_idx_s[w]; }
      P.best_out[2 * blockIdx.x] = bd / H;
      P.best_out[2 * blockIdx.x + 1] = bd % H;
    }
  }
}

// per-locus total LL = sum of the sample normalisers in sample order (genotyper.cpp:72-78)
__global__ void k_total_ll(const PostParams P) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= P.n_loci) return;
  double t = 0.0;
  for (int s = P.locus_sample_off[l]; s < P.locus_sample_off[l + 1]; s++) t += P.sample_ll_out[s];
  P.total_ll_out[l] = t;
}

cudaError_t launch_posteriors(const PostParams& p, cudaStream_t stream) {
  if (p.n_samples <= 0) return cudaSuccess;
  k_posteriors<<<p.n_samples, POST_THREADS, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.total_ll_out) {
    k_total_ll<<<(p.n_loci + 127) / 128, 128, 0, stream>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

// ------------------------------------------------------------------------------------------
// K4: EM stutter learner.  One CTA per locus runs EMStutterGenotyper::train
// (em_stutter_genotyper.cpp:170-226) to convergence: E-step = length-only alignment
// probabilities (a log_stutter_pmf table) + sample posteriors + read phase posteriors (recomputed
// on the fly, never materialised: the reference's R*A*A*2 array is the largest object of the
// learner), M-step = allele frequencies + the 7 approximate log-sum-exp buckets of the stutter
// parameters.  Exact exp/log come from CUDA's libm (ulp-level differences from glibc); the
// approximate log-sum-exps are the bit-faithful replicas of fastapprox.cuh.
// ------------------------------------------------------------------------------------------
#define EM_THREADS 512
#define EM_WARPS (EM_THREADS / 32)
#define EM_NEG_MAX (-1.7976931348623157e308)

struct EmModel {   // StutterModel log parameters (stutter_model.h:43-58)
  double in_step, in_nostep, in_up, in_down, out_step, out_nostep, out_up, out_down, equal;
};
__device__ __forceinline__ double em_pmf(const EmModel& m, int period, int sample_bps, int read_bps) {   // stutter_model.cpp:29-53
  const int d = read_bps - sample_bps;
  if (d % period != 0) {
    const int eff = d - d / period;
    return eff < 0 ? m.out_down + m.out_nostep + m.out_step * (-eff - 1) : m.out_up + m.out_nostep + m.out_step * (eff - 1);
  }
  const int reps = d / period;
  if (reps == 0) return m.equal;
  return reps < 0 ? m.in_down + m.in_nostep + m.in_step * (-reps - 1) : m.in_up + m.in_nostep + m.in_step * (reps - 1);
}
__device__ __forceinline__ double exact_lse2(double a, double b) {   // mathops.cpp:51-56
  return a > b ? a + log(1 + exp(b - a)) : b + log(1 + exp(a - b));
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// K3b: genotypes and likelihoods from the posteriors: Genotyper::extract_genotypes_and_likelihoods
// (genotyper.cpp:129-251) + calc_PLs (:99-104) + calc_gl_diff (:106-127).  One warp per (locus,
// sample).  The haplotype -> allele marginalisation is an exact log-sum-exp (the reference streams
// it; same value up to libm ulps); GL averaging uses the approximate two-argument form like the
// reference.
// ------------------------------------------------------------------------------------------
#define LOG_E_BASE_10 0.4342944819   /* mathops.cpp:11, as written there */

__global__ void __launch_bounds__(128) k_extract(const ExtractParams P) {
  const int lane = threadIdx.x & 31;
  const int sidx = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (sidx >= P.n_samples) return;
  const ExtractSample S = P.samples[sidx];
  const int H = S.n_haps, V = S.n_variants;
  const bool hap1 = S.haploid != 0;
  const int G = hap1 ? V : V * (V + 1) / 2, PG = hap1 ? V : V * V;
  const double* sp = P.post + S.post_off;
  const int32_t* h2a = P.hap_to_allele + S.h2a_off;
  double* gl = P.gl + S.gl_off;
  double* pgl = P.phased_gl + S.pgl_off;
  int32_t* pl = P.pl + S.gl_off;
  const double sll = P.sample_ll[sidx];

  // get_optimal_haplotypes (:82-97): first maximum
  double best = -1.7976931348623157e308;
  int best_d = H * H;
  for (int d = lane; d < H * H; d += 32) {
    const double v = sp[d];
    if (v > best) { best = v; best_d = d; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(FULL, best, o);
    const int od = __shfl_xor_sync(FULL, best_d, o);
    if (ov > best || (ov == best && od < best_d)) { best = ov; best_d = od; }
  }
  const int ba = best_d / H, bb = best_d % H;
  const int ga = h2a[ba], gb = h2a[bb];

  // marginalise haplotype pairs to allele pairs (:152-171).  totals live in a per-sample scratch:
  // the tail of the phased-GL slot when diploid (it has exactly V*V entries), else computed twice.
  // For the haploid case only the diagonal is ever used.
  auto total_of = [&](int va, int vb) {
    double mx = -1.7976931348623157e308 / 2;
    for (int a = 0; a < H; a++) {
      if (h2a[a] != va) continue;
      for (int b = 0; b < H; b++) if (h2a[b] == vb) mx = dmax(mx, sp[a * H + b]);
    }
    double sum = 0.0;
    for (int a = 0; a < H; a++) {
      if (h2a[a] != va) continue;
      for (int b = 0; b < H; b++) if (h2a[b] == vb) sum += exp(sp[a * H + b] - mx);
    }
    return mx + log(sum);
  };
  const double hom = hap1 ? -P.int_logs[H] : P.int_logs[2] - P.int_logs[H] - P.int_logs[H + 1];
  const double het = hap1 ? 0.0 : -P.int_logs[H] - P.int_logs[H + 1];
  const double gl_nconfig = hap1 ? P.int_logs[2] + P.int_logs[H] - P.int_logs[V] : P.int_logs[2] + 2 * (P.int_logs[H] - P.int_logs[V]);
  const double pgl_nconfig = hap1 ? P.int_logs[H] - P.int_logs[V] : 2 * (P.int_logs[H] - P.int_logs[V]);

  if (!hap1) {
    for (int g = lane; g < V * V; g += 32) pgl[g] = total_of(g / V, g % V);   // raw totals first
    __syncwarp();
    if (lane == 0) {
      const double lp = pgl[V * ga + gb];
      P.log_phased[sidx] = lp;
      P.log_unphased[sidx] = ga == gb ? lp : exact_lse2(lp, pgl[V * gb + ga]);
    }
    for (int i = lane; i < G; i += 32) {   // i -> (i1 >= i2)
      int i1 = (int)((sqrt(8.0 * i + 1.0) - 1.0) / 2.0);
      while ((i1 + 1) * (i1 + 2) / 2 <= i) i1++;
      while (i1 * (i1 + 1) / 2 > i) i1--;
      const int i2 = i - i1 * (i1 + 1) / 2;
      const double corr = (i1 == i2 ? hom : het) + gl_nconfig;
      gl[i] = (sll - corr + lse2(pgl[i1 * V + i2], pgl[i2 * V + i1])) * LOG_E_BASE_10;
    }
    __syncwarp();
    for (int g = lane; g < V * V; g += 32) {
      const double corr = ((g / V) == (g % V) ? hom : het) + pgl_nconfig;
      pgl[g] = (sll - corr + pgl[g]) * LOG_E_BASE_10;
    }
  } else {
    for (int v = lane; v < V; v += 32) {
      const double tot = total_of(v, v);
      if (v == ga) { P.log_phased[sidx] = tot; P.log_unphased[sidx] = tot; }   // haploid: ga == gb
      gl[v] = (sll - (hom + gl_nconfig) + lse2(tot, tot)) * LOG_E_BASE_10;
      pgl[v] = (sll - (hom + pgl_nconfig) + tot) * LOG_E_BASE_10;
    }
    if (ga != gb && lane == 0) {   // cannot happen with the haploid priors, kept for arbitrary inputs
      const double lp = total_of(ga, gb);
      P.log_phased[sidx] = lp;
      P.log_unphased[sidx] = exact_lse2(lp, total_of(gb, ga));
    }
  }
  __syncwarp();
  if (lane == 0) {
    P.best_hap[2 * sidx] = ba; P.best_hap[2 * sidx + 1] = bb;
    P.best_gt[2 * sidx] = ga; P.best_gt[2 * sidx + 1] = gb;
    P.hap_log_phased[sidx] = sp[ba * H + bb];
    P.hap_log_unphased[sidx] = ba != bb ? lse2(sp[ba * H + bb], sp[bb * H + ba]) : sp[ba * H + bb];
  }
  // GLDIFF (:106-127) and PLs (:99-104)
  double mg = -1.7976931348623157e308;
  for (int i = lane; i < G; i += 32) mg = dmax(mg, gl[i]);
  mg = warp_max(mg);
  double second = -1.7976931348623157e308;
  for (int i = lane; i < G; i += 32) if (gl[i] < mg) second = dmax(second, gl[i]);
  second = warp_max(second);
  if (second == -1.7976931348623157e308) second = mg;
  for (int i = lane; i < G; i += 32) {
    const int v = (int)(-10 * (gl[i] - mg));
    pl[i] = v < 999 ? v : 999;
  }
  if (lane == 0) {
    double d;
    if (H == 1) d = -1000;
    else {
      const int hi = ga > gb ? ga : gb, lo = ga > gb ? gb : ga;
      const int idx = hap1 ? ga : hi * (hi + 1) / 2 + lo;
      d = fabs(mg - gl[idx]) < 1e-10 ? mg - second : gl[idx] - mg;
    }
    P.gl_diff[sidx] = d;
  }
}

cudaError_t launch_extract(const ExtractParams& p, cudaStream_t stream) {
  if (p.n_samples <= 0) return cudaSuccess;
  k_extract<<<(p.n_samples + 3) / 4, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(EM_THREADS) k_em_train(const EmParams P) {
  extern __shared__ __align__(16) double em_smem[];
  const EmLocus L = P.loci[blockIdx.x];
  const int A = L.n_alleles, S = L.n_samples, R = L.n_reads, AA = A * A;
  double* s_T = em_smem;                       // [A][A]: log_stutter_pmf(bps[a], bps[c])
  double* s_prior = s_T + AA;                  // [A]
  double* s_newprior = s_prior + A;            // [A]
  double* s_red = s_newprior + A;              // [7][EM_WARPS]
  double* s_bucket = s_red + 7 * EM_WARPS;     // [7]
  __shared__ EmModel s_model;
  __shared__ double s_prm[6];
  __shared__ double s_LL, s_newLL;
  __shared__ int s_state;                      // 0 = continue, 1 = converged
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* bps = P.bps + L.allele_off;
  double* post = P.post + L.post_off;
  double* rowlse = P.rowlse + L.row_off;
  double* sll = P.sample_ll + L.sample0;
  const int32_t* allele_of = P.allele_of + L.read0;
  const int32_t* label = P.sample_label + L.read0;
  const double* p1 = P.log_p1 + L.read0;
  const double* p2 = P.log_p2 + L.read0;
  const double LH = P.log_one_half;

  for (int a = tid; a < A; a += EM_THREADS) s_prior[a] = P.gt_prior[L.allele_off + a];
  if (tid < 6) s_prm[tid] = P.params[6 * (size_t)blockIdx.x + tid];
  if (tid == 0) { s_LL = EM_NEG_MAX; s_state = 0; }
  __syncthreads();
  int iter = 1;
  bool converged = false;
  while (iter <= P.max_iter) {
    if (tid == 0) {
      EmModel m;
      m.in_step = log(1 - s_prm[0]); m.in_nostep = log(s_prm[0]); m.in_up = log(s_prm[1]); m.in_down = log(s_prm[2]);
      m.out_step = log(1 - s_prm[3]); m.out_nostep = log(s_prm[3]); m.out_up = log(s_prm[4]); m.out_down = log(s_prm[5]);
      m.equal = log(1 - s_prm[1] - s_prm[2] - s_prm[4] - s_prm[5]);
      s_model = m;
    }
    __syncthreads();
    for (int d = tid; d < AA; d += EM_THREADS) s_T[d] = em_pmf(s_model, L.period, bps[d / A], bps[d % A]);
    __syncthreads();

    // ---- E-step: sample posteriors (genotyper.cpp:44-80 with the allele-frequency priors of :129-144) ----
    for (int s = warp; s < S; s += EM_WARPS) {
      const int r0 = P.sample_read_off[L.sample0 + s] - L.read0, r1 = P.sample_read_off[L.sample0 + s + 1] - L.read0;
      double* sp = post + (size_t)s * AA;
      double mx = EM_NEG_MAX;
      for (int d = lane; d < AA; d += 32) {
        const int a = d / A, b = d % A;
        double acc = L.haploid ? (a == b ? s_prior[a] : EM_NEG_MAX / 2) : s_prior[a] + s_prior[b];
        for (int r = r0; r < r1; r++) {
          const int c = allele_of[r];
          acc += lse2(LH + p1[r] + s_T[a * A + c], LH + p2[r] + s_T[b * A + c]);
        }
        sp[d] = acc;
        mx = dmax(mx, acc);
      }
      mx = warp_max(mx);
      double sum = 0.0;
      for (int d = lane; d < AA; d += 32) sum += exp(sp[d] - mx);
      sum = warp_sum(sum);
      const double total = mx + log(sum);
      for (int d = lane; d < AA; d += 32) sp[d] -= total;
      if (lane == 0) sll[s] = total;
    }
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int s = 0; s < S; s++) t += sll[s];
      s_newLL = t;
      if (t < s_LL + 1e-10) { s_state = 1; s_LL = t; }   // em_stutter_genotyper.cpp:195-199
    }
    __syncthreads();
    if (s_state) { converged = true; break; }

    // ---- M-step 1: allele frequencies (recalc_log_gt_priors, :21-56) ----
    for (int i = tid; i < S * A; i += EM_THREADS) {   // log_sum_exp over the second allele
      const double* row = post + (size_t)i * A;
      double mx = row[0];
      for (int b = 1; b < A; b++) mx = dmax(mx, row[b]);
      double sum = 0.0;
      for (int b = 0; b < A; b++) sum += exp(row[b] - mx);
      rowlse[i] = mx + log(sum);
    }
    __syncthreads();
    for (int b = warp; b < A; b += EM_WARPS) {
      double mx = EM_NEG_MAX / 2;
      for (int s = lane; s < S; s += 32) mx = dmax(mx, rowlse[(size_t)s * A + b]);
      for (int i = lane; i < S * A; i += 32) mx = dmax(mx, post[(size_t)i * A + b]);
      mx = warp_max(mx);
      double sum = 0.0;
      for (int s = lane; s < S; s += 32) sum += exp(rowlse[(size_t)s * A + b] - mx);
      for (int i = lane; i < S * A; i += 32) sum += exp(post[(size_t)i * A + b] - mx);
      sum = warp_sum(sum);
      if (lane == 0) s_newprior[b] = mx + log(sum);
    }
    __syncthreads();

    // ---- M-step 2: stutter parameters (recalc_stutter_model, :63-127; phase posteriors :152-168) ----
    // buckets: 0 in_up 1 in_down 2 in_eq 3 in_diffs 4 out_up 5 out_down 6 out_diffs
    for (int pass = 0; pass < 2; pass++) {
      double acc[7];
      const double log11 = log(1.1);
      if (pass == 0) {
#pragma unroll
        for (int k = 0; k < 7; k++) acc[k] = tid == 0 ? 0.0 : EM_NEG_MAX;       // pseudocount terms (:67-70)
        if (tid == 0) { acc[3] = dmax(0.0, log11); acc[6] = dmax(0.0, log11); }
      } else {
#pragma unroll
        for (int k = 0; k < 7; k++) acc[k] = tid == 0 ? lse_term(0.0, s_bucket[k]) : 0.0;
        if (tid == 0) { acc[3] += lse_term(log11, s_bucket[3]); acc[6] += lse_term(log11, s_bucket[6]); }
      }
      const long long total_items = (long long)R * AA;
      for (long long it = tid; it < total_items; it += EM_THREADS) {
        const int r = (int)(it / AA), d = (int)(it % AA);
        const int a = d / A, b = d % A;
        const int c = allele_of[r];
        const double one = LH + p1[r] + s_T[a * A + c], two = LH + p2[r] + s_T[b * A + c];
        const double both = lse2(one, two);
        const double g = post[(size_t)label[r] * AA + d];
        const int rb = bps[c];
#pragma unroll
        for (int ph = 0; ph < 2; ph++) {
          const double f = g + ((ph == 0 ? one : two) - both);
          const int diff = rb - bps[ph == 0 ? a : b];
          int k_main, k_diff = -1, eff = 0;
          if (diff == 0) k_main = 2;
          else if (diff % L.period != 0) { eff = diff - diff / L.period; k_main = diff > 0 ? 4 : 5; k_diff = 6; }
          else { eff = diff / L.period; k_main = diff > 0 ? 0 : 1; k_diff = 3; }
          const double fd = k_diff >= 0 ? f + P.int_logs[eff < 0 ? -eff : eff] : 0.0;
#pragma unroll
          for (int k = 0; k < 7; k++) {
            if (k == k_main) acc[k] = pass ? acc[k] + lse_term(f, s_bucket[k]) : dmax(acc[k], f);
            if (k == k_diff) acc[k] = pass ? acc[k] + lse_term(fd, s_bucket[k]) : dmax(acc[k], fd);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 7; k++) {
        const double v = pass ? warp_sum(acc[k]) : warp_max(acc[k]);
        if (lane == 0) s_red[k * EM_WARPS + warp] = v;
      }
      __syncthreads();
      if (tid < 7) {
        double v = s_red[tid * EM_WARPS];
        for (int w = 1; w < EM_WARPS; w++) v = pass ? v + s_red[tid * EM_WARPS + w] : dmax(v, s_red[tid * EM_WARPS + w]);
        s_bucket[tid] = pass ? lse_finish(s_bucket[tid], v) : v;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double iu = s_bucket[0], id = s_bucket[1], ie = s_bucket[2], idf = s_bucket[3];
      const double ou = s_bucket[4], od = s_bucket[5], odf = s_bucket[6];
      const double ot = lse2(ou, od);
      double prm[6];
      prm[0] = fmin(0.999, exp(exact_lse2(iu, id) - idf));
      prm[3] = fmin(0.999, exp(ot - odf));
      const double m3 = dmax(dmax(iu, id), ie);
      const double in_all = m3 + log(exp(iu - m3) + exp(id - m3) + exp(ie - m3));   // mathops.cpp:58-61
      const double lt = exact_lse2(in_all, ot);
      prm[1] = exp(iu - lt); prm[2] = exp(id - lt); prm[4] = exp(ou - lt); prm[5] = exp(od - lt);
      bool close = true;
      for (int k = 0; k < 6; k++) { close = close && fabs(s_prm[k] - prm[k]) < 0.0001; s_prm[k] = prm[k]; }
      const double abs_change = s_newLL - s_LL, frac_change = -(s_newLL - s_LL) / s_LL;
      s_LL = s_newLL;
      if ((abs_change < P.min_abs && frac_change < P.min_frac) || close) s_state = 1;
      // normalise the new allele frequencies (:50-55)
      double mx = s_newprior[0];
      for (int a = 1; a < A; a++) mx = dmax(mx, s_newprior[a]);
      double sum = 0.0;
      for (int a = 0; a < A; a++) sum += exp(s_newprior[a] - mx);
      const double total = mx + log(sum);
      for (int a = 0; a < A; a++) s_prior[a] = s_newprior[a] - total;
    }
    __syncthreads();
    if (s_state) { converged = true; break; }
    iter++;
  }
  if (tid < 6) P.params[6 * (size_t)blockIdx.x + tid] = s_prm[tid];
  for (int a = tid; a < A; a += EM_THREADS) P.gt_prior[L.allele_off + a] = s_prior[a];
  if (tid == 0) {
    P.converged[blockIdx.x] = converged;
    P.iters[blockIdx.x] = iter < P.max_iter ? iter : P.max_iter;
    P.final_ll[blockIdx.x] = s_LL;
  }
}

cudaError_t launch_em(const EmParams& p, int max_alleles, cudaStream_t stream) {
  if (p.n_loci <= 0) return cudaSuccess;
  const size_t smem = ((size_t)max_alleles * max_alleles + 2 * (size_t)max_alleles + 7 * EM_WARPS + 7) * sizeof(double);
  cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void*>(&k_em_train));
  if (e != cudaSuccess) return e;
  k_em_train<<<p.n_loci, EM_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace hipstr
