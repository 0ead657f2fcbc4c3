/*
 * trace.cu -- K5: alignment traceback (HapAligner::trace_optimal_aln, SeqAlignment/HapAligner.cpp:711-722
 * -> process_read(retrace_aln = true) :636-690 -> retrace :363-571).
 *
 * ONE THREAD per (pooled read, haplotype) trace.  The reference keeps the three FULL matrices of both
 * sides of the seed and, walking back, re-derives at every visited cell which predecessor was best
 * (within TRACE_LL_TOL, preferring different states on the two sides).  Those choices depend only on
 * values that are all at hand when the cell is computed in the forward pass, so here they are TAKEN
 * in the forward pass and stored as one byte per cell (2 bits for the match state, 1 each for the
 * insertion and deletion states); the matrices themselves shrink to two rolling rows, the last
 * column of every row (for the seed placement) and the best artifact size / position of every
 * repeat-block column.  A trace needs ~25 KB instead of ~400 KB, all of it L1/L2-resident, and the
 * strictly sequential walk back reads bytes.  Recurrences run in the reference's order (bit-identical
 * cells, hence identical tie decisions).
 * The repeat-block evaluator replays the same host-unrolled programs as K1 (layout.h); the position
 * of the best artifact is read off the program: after a step, the reference's position counter is
 * one above the next step's position.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/hipstr_b200.h"
#include "fastapprox.cuh"
#include "kernels.h"
#include "layout.h"

namespace hipstr {

#define T_IMPOSSIBLE (-1000000000.0)
#define T_INS_TO_INS (-1.0)
#define T_INS_TO_MATCH (-0.4586751453870818910216436)
#define T_DEL_TO_DEL (-1.0)
#define T_DEL_TO_MATCH (-0.4586751453870818910216436)
#define T_TRACE_TOL 0.001                    /* HapAligner.cpp:345 */
#define T_MIN_SNP_LOG_CORRECT (-0.0043648054) /* HapAligner.cpp:24 */

__device__ __forceinline__ double tmax(double a, double b) { return a > b ? a : b; }


/* A per-thread array whose element i lives at p[i * 32]: the 32 lanes of a warp interleave their slabs, and because
 * all lanes index their matrices with the SAME row pitch (the warp's longest read side), lanes that run the DP in
 * lockstep touch one 256-byte line per access instead of 32 sectors 0.4 MB apart. */
template <class T>
struct Lane {
  T* p;
  __device__ __forceinline__ T& operator[](long i) const { return p[i * 32]; }
  __device__ __forceinline__ Lane operator+(long off) const { return Lane{p + off * 32}; }
};
typedef Lane<double> Mat;
typedef Lane<int> IMat;
typedef Lane<unsigned char> BMat;

/* Two rolling rows of the three matrices. */
struct Rows {
  Mat M[2], I[2], D[2];
};

// One side of the seed: column k of the side is read base (rev ? n_read-1-k : k).
struct TSide {
  const uint8_t* bases;   // codes, read order
  Lane<double> E;         // per-thread emission table [read position][6]: log P(read base | haplotype code 0..4), then
                          // log P(base correct); built once per trace so that an emission is ONE load instead of the
                          // dependent chain quality byte -> base byte -> quality table
  int n, rev, n_read;
  int pitch;              // row pitch of this side's matrices (>= n, uniform across the warp)
  __device__ __forceinline__ int ridx(int k) const { return rev ? n_read - 1 - k : k; }
  __device__ __forceinline__ int code(int k) const { return bases[ridx(k)]; }
  __device__ __forceinline__ double lc(int k) const { return E[ridx(k) * 6 + 5]; }
  __device__ __forceinline__ double emit(int k, int x) const { return E[ridx(k) * 6 + x]; }
};

struct TRep {
  const uint8_t* s;          // oriented allele codes
  const DevProgEntry* progs;
  const double* logrun;
  const DevRep* rep;
  const double* int_logs;
  int B, p, left_align;
};

// match_probs_[q] (StutterAlignerClass.cpp:12-53), recomputed on demand
__device__ double t_match(const TSide& sd, const TRep& r, int q) {
  const int terms = min(q + 1, r.B);
  double acc = 0.0;
  for (int t = 0; t < terms; t++) acc += sd.emit(q - t, r.s[r.B - 1 - t]);
  return acc;
}

// Replays a position walk and tracks the reference's best position (StutterAlignerClass.cpp:92-95,137-140: the
// position counter after a step is one above the next step's position, so best_pos = 1 - i == -next.pos).  The terms
// of the walk's log-sum-exp are parked in a small per-thread array so that maximum and sum need ONE replay (the float
// summands add exactly in a double, in any order); a walk with more terms than slots is replayed for the sum.
#define T_WALK_SLOTS 16
template <bool INS>
__device__ double t_walk(const TSide& sd, const TRep& r, int prog_index, int stop, int j, int units, double lp0,
                         int tail_base, int& best_pos) {
  const int CB = HIPSTR_VAL_STRIDE * 8;
  double mx = lp0, best = lp0;
  best_pos = 0;
  double terms[T_WALK_SLOTS];
  int n = 0;
  const DevProgEntry* e = r.progs + prog_index;
  const double* lr = r.logrun + prog_index;
  double lp = lp0;
  while (e->pos > stop) {
    if (e->moves) {
      // offsets are pos * CB + code * 8 (see DevProgEntry); recover the two base codes
      const int xa = (e->off_a - e->pos * CB) / 8, xb = (e->off_b - e->pos * CB) / 8;
      if (INS) {
        for (int m = 0; m < units; m++) {
          const int col = j - r.p + e->pos - m * r.p;
          lp -= sd.emit(col, xa);
          lp += sd.emit(col, xb);
        }
      } else {
        lp -= sd.emit(j + e->pos, xa);
        lp += sd.emit(j + e->pos, xb);
      }
    }
    const double term = lp + *lr;
    if (n < T_WALK_SLOTS) terms[n] = term;
    n++;
    mx = tmax(mx, term);
    if (lp > best || (r.left_align && lp == best)) { best_pos = -(e + 1)->pos; best = lp; }
    e++; lr++;
  }
  const int fin = e->pos;
  const bool has_tail = INS ? (fin > -tail_base) : (-fin < tail_base);
  double tail = 0.0;
  if (has_tail) {
    tail = __ldg(r.int_logs + (tail_base + fin)) + lp;
    mx = tmax(mx, tail);
  }
  double total = lse_term(lp0, mx);
  if (n <= T_WALK_SLOTS) {
    for (int k = 0; k < n; k++) total += lse_term(terms[k], mx);
  } else {   // rare: replay the walk, summing against the known maximum
    e = r.progs + prog_index;
    lr = r.logrun + prog_index;
    lp = lp0;
    while (e->pos > stop) {
      if (e->moves) {
        const int xa = (e->off_a - e->pos * CB) / 8, xb = (e->off_b - e->pos * CB) / 8;
        if (INS) {
          for (int m = 0; m < units; m++) {
            const int col = j - r.p + e->pos - m * r.p;
            lp -= sd.emit(col, xa);
            lp += sd.emit(col, xb);
          }
        } else {
          lp -= sd.emit(j + e->pos, xa);
          lp += sd.emit(j + e->pos, xb);
        }
      }
      total += lse_term(lp + *lr, mx);
      e++; lr++;
    }
  }
  if (has_tail) total += lse_term(tail, mx);
  return lse_finish(mx, total);
}

__device__ __forceinline__ int t_pick3(bool rev, double v1, double v2, double v3) {   // HapAligner.cpp:346-358
  if (!rev) {
    if (v1 > v2 + T_TRACE_TOL) return v1 > v3 + T_TRACE_TOL ? 0 : 2;
    return v2 > v3 + T_TRACE_TOL ? 1 : 2;
  }
  if (v3 > v2 + T_TRACE_TOL) return v3 > v1 + T_TRACE_TOL ? 2 : 0;
  return v2 > v1 + T_TRACE_TOL ? 1 : 0;
}
__device__ __forceinline__ int t_pick2(bool rev, double v1, double v2) {              // :360-361
  if (!rev) return v1 > v2 + T_TRACE_TOL ? 0 : 1;
  return v2 > v1 + T_TRACE_TOL ? 1 : 0;
}


// Repeat block of one side: the super-row `out_row` from the row above it (HapAligner.cpp:62-109).
__device__ void t_repeat_block(const TSide& sd, const TRep& r, const Mat M_prev, const Mat M_out, const Mat I_out,
                               const Mat D_out, const Mat match, const IMat art_size, const IMat art_pos) {
  const int B = r.B, p = r.p, n = sd.n;
  // match_probs_ of load_read, once per (side, allele) like the reference's table instead of once per use
  for (int q = 0; q < n; q++) match[q] = t_match(sd, r, q);
  for (int j = 0; j < n; j++) {
    double probs[HIPSTR_NUM_ARTIFACTS];
    double best = T_IMPOSSIBLE;
    art_size[j] = -10000;
    art_pos[j] = 0;
    for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) {
      const int units = a - HIPSTR_MAX_ARTIFACT_UNITS, D = units * p;
      const int base_len = min(B + D, j + 1);
      int pos = -1;
      double v = T_IMPOSSIBLE;
      if (base_len >= 0) {
        double pr;
        if (units == 0) pr = match[j];
        else if (units < 0) {
          const int k = -units;
          double lp0 = -__ldg(r.int_logs + (B + D + 1));
          const int q = j - D;
          if (q <= n - 1) {
            double pre = 0.0;
            for (int t = 0; t < -D; t++) pre += sd.emit(q - t, r.s[B - 1 - t]);
            lp0 += match[q] - pre;
          } else {
            for (int t = 0; t < base_len; t++) lp0 += sd.emit(j - t, r.s[B - 1 - t + D]);
          }
          pr = t_walk<false>(sd, r, __ldg(r.rep->prog_off + k), -base_len, j, k, lp0, B + D, pos);
        } else {
          double ins = 0.0;
          const int upto = min(D, j + 1);
          for (int t = 0; t < upto; t++) {
            const int m = t % p;
            ins += (m < B) ? sd.emit(j - t, r.s[B - 1 - m]) : sd.lc(j - t);
          }
          double lp0 = -__ldg(r.int_logs + (B + 1)) + ins;
          lp0 += (base_len > D) ? match[j - D] : 0.0;
          const int stop = -min(max(0, base_len - D), B);
          pr = t_walk<true>(sd, r, __ldg(r.rep->prog_off), stop, j, units, lp0, B, pos);
        }
        const double pre_row = (j - base_len < 0) ? 0.0 : M_prev[j - base_len];
        v = __ldg(r.rep->art + a) + pr + pre_row;
      }
      probs[a] = v;
      if (v > best) { art_size[j] = D; art_pos[j] = pos; best = v; }
    }
    double mx = probs[0];
    for (int a = 1; a < HIPSTR_NUM_ARTIFACTS; a++) mx = tmax(mx, probs[a]);
    double total = 0.0;
    for (int a = 0; a < HIPSTR_NUM_ARTIFACTS; a++) total += lse_term(probs[a], mx);
    M_out[j] = lse_finish(mx, total);
    I_out[j] = T_IMPOSSIBLE;
    D_out[j] = T_IMPOSSIBLE;
  }
}

// One side of align_seq_to_hap (HapAligner.cpp:26-161) with two rolling rows.  For every cell the predecessor choices
// of retrace (:363-571) are taken here and stored in `dec` ([row][pitch]): bits 0-1 = best of (insertion to the
// left, deletion on the diagonal, match on the diagonal) for the match state, bit 2 = deletion state came from a
// match, bit 3 = insertion state came from a match.  lastcol[row] = M[row][n-1] (compute_aln_logprob reads those).
__device__ __forceinline__ double t_fill_side(const TraceParams& P, const TSide& sd, const DevHapSide& hs, const Rows& rw, const BMat dec,
                              const Mat lastcol, const Mat match, const IMat art_size, const IMat art_pos) {
  const int n = sd.n;
  const long pitch = sd.pitch;
  const bool rev = sd.rev != 0;
  const uint8_t* seq = P.hapbytes + hs.seq_off;
  const uint8_t* rows = P.hapbytes + hs.row_off;
  int cur = 0;   // rw.*[cur] holds the row computed last
  double run = 0.0;
  for (int j = 0; j < n; j++) {
    rw.M[0][j] = sd.emit(j, seq[0]) + run;
    rw.I[0][j] = sd.lc(j) + run;
    rw.D[0][j] = T_IMPOSSIBLE;
    run += sd.lc(j);
  }
  if (n > 0) lastcol[0] = rw.M[0][n - 1];
  for (int b = 0; b < hs.n_blocks; b++) {
    const DevBlock blk = P.blocks[hs.blk_off + b];
    if (blk.rep >= 0) {
      const DevRep* rep = P.reps + blk.rep;
      TRep r;
      r.s = P.hapbytes + rep->seq_off; r.progs = P.progs; r.logrun = P.prog_logrun; r.rep = rep; r.int_logs = P.int_logs;
      r.B = rep->len; r.p = rep->period; r.left_align = rep->left_align;
      const int nxt = cur ^ 1;
      t_repeat_block(sd, r, rw.M[cur], rw.M[nxt], rw.I[nxt], rw.D[nxt], match, art_size + pitch * b, art_pos + pitch * b);
      cur = nxt;
      if (n > 0) lastcol[blk.row_start + blk.len - 1] = rw.M[cur][n - 1];
      continue;
    }
    for (int row = blk.row_start + (b == 0 ? 1 : 0); row < blk.row_start + blk.len; row++) {
      const int hc = seq[row];
      const int info = rows[row], hp = info & 15;
      const bool after = (info & HIPSTR_ROW_AFTER_REPEAT) != 0;
      const double m2m = __ldg(P.trans + hp), m2i = __ldg(P.trans + 16 + hp), m2d = __ldg(P.trans + 32 + hp);
      const Mat pM = rw.M[cur], pD = rw.D[cur];
      const int nxt = cur ^ 1;
      const Mat cM = rw.M[nxt], cI = rw.I[nxt], cD = rw.D[nxt];
      const BMat drow = dec + pitch * row;
      if (n > 0) {
        const double upM = pM[0], upD = pD[0];
        cM[0] = sd.emit(0, hc);
        cI[0] = after ? T_IMPOSSIBLE : sd.lc(0);
        cD[0] = after ? T_IMPOSSIBLE : tmax(upD + T_DEL_TO_DEL, upM + T_DEL_TO_MATCH);
        drow[0] = (unsigned char)(t_pick2(rev, upD + T_DEL_TO_DEL, upM + T_DEL_TO_MATCH) << 2);
      }
      double diagM = n > 0 ? pM[0] : 0.0, diagD = n > 0 ? pD[0] : 0.0, leftI = n > 0 ? cI[0] : 0.0;
      for (int j = 1; j < n; j++) {
        const double e = sd.emit(j, hc);
        const double upM = pM[j], upD = pD[j];
        double m, i, d;
        if (after) {
          m = e + diagM;
          i = T_IMPOSSIBLE;
          d = T_IMPOSSIBLE;
        } else {
          m = e + tmax(leftI + m2i, tmax(diagM + m2m, diagD + m2d));
          i = sd.lc(j) + tmax(diagM + T_INS_TO_MATCH, leftI + T_INS_TO_INS);
          d = tmax(upM + T_DEL_TO_MATCH, upD + T_DEL_TO_DEL);
        }
        // the choices retrace would make standing on this cell, from the very values it would read
        drow[j] = (unsigned char)(t_pick3(rev, leftI + m2i, diagD + m2d, diagM + m2m) |
                                  (t_pick2(rev, upD + T_DEL_TO_DEL, upM + T_DEL_TO_MATCH) << 2) |
                                  (t_pick2(rev, leftI + T_INS_TO_INS, diagM + T_INS_TO_MATCH) << 3));
        cM[j] = m; cI[j] = i; cD[j] = d;
        diagM = upM; diagD = upD; leftI = i;
      }
      cur = nxt;
      if (n > 0) lastcol[row] = rw.M[cur][n - 1];
    }
  }
  return run;
}

struct TAcc {
  int32_t* stutter; int32_t* lo; int32_t* hi;   // per forward block
  int32_t* indels; int32_t* snps;
  int n_indels, n_snps, ins, del;
  __device__ void touch(int block, int read_index) {
    if (read_index < lo[block]) lo[block] = read_index;
    if (read_index > hi[block]) hi[block] = read_index;
  }
  __device__ void indel(int pos, int size) {
    if (n_indels < HIPSTR_MAX_TRACE_INDELS) { indels[2 * n_indels] = pos; indels[2 * n_indels + 1] = size; }
    n_indels++;
  }
  __device__ void snp(int pos, int base) {
    if (n_snps < HIPSTR_MAX_TRACE_SNPS) { snps[2 * n_snps] = pos; snps[2 * n_snps + 1] = base; }
    n_snps++;
  }
};

// Homopolymer class of a flank row: the lowering stored min(15, max(h(i), h(i-1))) per row.
// retrace (HapAligner.cpp:363-571) for one side; writes ops BACKWARDS-in-walk order into `ops`
// (the caller reverses the left side); returns the number of ops.
__device__ int t_walk_back(const TraceParams& P, const TSide& sd, const DevHapSide& hs, const int32_t* start_of,
                           const BMat dec, const IMat art_size, const IMat art_pos,
                           int block_index, int base_index, long matrix_index, char* ops, TAcc& acc) {
  const int n = sd.n, nb = hs.n_blocks;
  const long pitch = sd.pitch;
  const bool rev = sd.rev != 0;
  const uint8_t* seq = P.hapbytes + hs.seq_off;
  const uint8_t* rows = P.hapbytes + hs.row_off;
  const char* letters = "ACTGN";   // base codes of the host lowering -> characters (flatten.cpp base_code)
  int seq_index = n - 1, type = 0, n_ops = 0;
  while (block_index >= 0) {
    const DevBlock blk = P.blocks[hs.blk_off + block_index];
    const int fw_block = rev ? nb - 1 - block_index : block_index;
    if (blk.rep >= 0) {
      const int size = art_size[pitch * block_index + seq_index], pos = art_pos[pitch * block_index + seq_index];
      const int len = blk.len;
      int i = 0;
      for (; i < min(seq_index + 1, pos); i++) { ops[n_ops++] = 'M'; acc.touch(fw_block, sd.ridx(seq_index - i)); }
      if (size < 0) for (int d = 0; d < -size; d++) ops[n_ops++] = 'D';
      else for (; i < min(seq_index + 1, pos + size); i++) { ops[n_ops++] = 'I'; acc.touch(fw_block, sd.ridx(seq_index - i)); }
      for (; i < min(len + size, seq_index + 1); i++) { ops[n_ops++] = 'M'; acc.touch(fw_block, sd.ridx(seq_index - i)); }
      acc.stutter[fw_block] = size;
      if (len + size >= seq_index + 1) return n_ops;
      matrix_index -= (len + size + pitch * len);
      type = 0;
      seq_index -= (len + size);
    } else {
      int prev_type = -1;
      int pos = start_of[block_index] + (rev ? -base_index : base_index);
      const int step = rev ? 1 : -1;
      int indel_seq_index = -1, indel_pos = -1;
      while (base_index >= 0 && seq_index >= 0) {
        if (type != prev_type) {
          if (prev_type == 1) { if (rev) acc.indel(indel_pos, indel_pos - pos); else acc.indel(pos + 1, pos - indel_pos); }
          else if (prev_type == 2) acc.indel(indel_pos + (rev ? 0 : 1), indel_seq_index - seq_index);
          if (type == 1 || type == 2) { indel_seq_index = seq_index; indel_pos = pos; }
          prev_type = type;
        }
        if (type == 0) {
          const int x = sd.code(seq_index);
          if (seq[blk.row_start + base_index] != x && sd.lc(seq_index) > T_MIN_SNP_LOG_CORRECT) acc.snp(pos, letters[x]);
          acc.touch(fw_block, sd.ridx(seq_index));
          ops[n_ops++] = 'M'; seq_index--; base_index--; pos += step;
        } else if (type == 1) {
          acc.del++; ops[n_ops++] = 'D'; base_index--; pos += step;
        } else {
          acc.ins++; acc.touch(fw_block, sd.ridx(seq_index)); ops[n_ops++] = 'I'; seq_index--;
        }
        if (seq_index == -1 || (base_index == -1 && block_index == 0)) {
          for (; seq_index != -1; seq_index--) ops[n_ops++] = 'S';
          return n_ops;
        }
        const int choice = dec[matrix_index];   // taken in the forward pass on this very cell
        if (type == 0) {
          const int best = choice & 3;
          if (best == 0) { type = 2; matrix_index -= 1; }
          else { type = best == 1 ? 1 : 0; matrix_index -= pitch + 1; }
        } else if (type == 1) {
          type = ((choice >> 2) & 1) == 0 ? 1 : 0;
          matrix_index -= pitch;
        } else {
          if (((choice >> 3) & 1) == 0) { type = 2; matrix_index -= 1; }
          else { type = 0; matrix_index -= pitch + 1; }
        }
      }
    }
    --block_index;
    if (block_index >= 0) base_index = P.blocks[hs.blk_off + block_index].len - 1;
  }
  return n_ops;
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(64, MIN_BLOCKS) k_trace(const TraceParams P) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_slots = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // whole warps iterate together (idle lanes of the last round still take part in the pitch reduction)
  for (int tr0 = slot - lane; tr0 < P.n_traces; tr0 += n_slots) {
    const bool live = tr0 + lane < P.n_traces;
    const int tr = live ? P.trace_order[tr0 + lane] : 0;   // traces sorted by (locus, haplotype, seed): lanes run in step
    int my_left = 0, my_right = 0;
    if (live) {
      const DevPool pl = P.pools[P.trace_pool[tr]];
      my_left = pl.seed;
      my_right = pl.len - pl.seed - 1;
    }
    const int pitchL = __reduce_max_sync(0xffffffffu, my_left), pitchR = __reduce_max_sync(0xffffffffu, my_right);
    if (!live) continue;
    const DevPool pool = P.pools[P.trace_pool[tr]];
    const int h = P.trace_hap[tr];
    const DevHapSide hsF = P.hapsides[pool.hap_rec0 + 2 * h];
    const DevHapSide hsR = P.hapsides[pool.hap_rec0 + 2 * h + 1];
    const int n = pool.len, seed = pool.seed, nL = seed, nR = n - seed - 1, hs_len = hsF.len, nb = hsF.n_blocks;
    // per-warp slab, lanes interleaved: two rolling rows of [M I D] (shared by the two sides, which run one after the
    // other), the last column of every row of both sides, the decision bytes of both sides, the artifact tables
    const Mat slab{P.slab + (size_t)(slot - lane) * P.slab_doubles + lane};
    const long pitch = max(pitchL, pitchR);
    Rows rw;
    rw.M[0] = slab; rw.M[1] = slab + pitch; rw.I[0] = slab + 2 * pitch; rw.I[1] = slab + 3 * pitch;
    rw.D[0] = slab + 4 * pitch; rw.D[1] = slab + 5 * pitch;
    const Mat lastL = slab + 6 * pitch, lastR = lastL + hs_len;
    const Mat matchq = lastR + hs_len;   // [pitch] match_probs_ of the repeat block being evaluated
    const Mat emis = matchq + pitch;     // [n][6]
    const BMat decL{P.dec_slab + (size_t)(slot - lane) * P.dec_bytes + lane};
    const BMat decR = decL + (long)pitchL * hs_len;
    const IMat arts{P.art_slab + (size_t)(slot - lane) * P.art_ints + lane};
    const IMat Ls = arts, Lp = Ls + (long)pitchL * nb, Rs = Lp + (long)pitchL * nb, Rp = Rs + (long)pitchR * nb;
    TSide L, R;
    L.bases = (const uint8_t*)P.bases + pool.seq_off;
    L.E = emis;
    {
      const uint8_t* quals = (const uint8_t*)P.quals + pool.seq_off;
      for (int r = 0; r < n; r++) {
        const double ok = __ldg(P.qual_lut + 2 * quals[r]), bad = __ldg(P.qual_lut + 2 * quals[r] + 1);
        const int b = L.bases[r];
        for (int x = 0; x < 5; x++) emis[r * 6 + x] = b == x ? ok : bad;
        emis[r * 6 + 5] = ok;
      }
    }
    L.n = nL; L.rev = 0; L.n_read = n; L.pitch = pitchL;
    R = L; R.n = nR; R.rev = 1; R.pitch = pitchR;
    // both sides through ONE inlined copy of the evaluator (a loop, not two call sites: half the code, and the
    // instruction cache of a divergent thread-per-trace kernel is a measured stall)
    double edge[2];
#pragma unroll 1
    for (int side = 0; side < 2; side++)
      edge[side] = t_fill_side(P, side ? R : L, side ? hsR : hsF, rw, side ? decR : decL, side ? lastR : lastL, matchq,
                               side ? Rs : Ls, side ? Rp : Lp);
    const double edgeL = edge[0], edgeR = edge[1];
    // best seed placement (compute_aln_logprob, HapAligner.cpp:163-231)
    const uint8_t* fseq = P.hapbytes + hsF.seq_off;
    const uint8_t* frow = P.hapbytes + hsF.row_off;
    const double prior = -__ldg(P.int_logs + hsF.n_seed_pos);
    const int sx = L.bases[seed];
    const double s_ok = emis[seed * 6 + 5], s_bad = emis[seed * 6 + (sx == 0 ? 1 : 0)];
    int max_index = 0;
    double best = prior + (sx == fseq[0] ? s_ok : s_bad) + edgeL + lastR[hs_len - 2];
    {
      const double v = prior + (sx == fseq[hs_len - 1] ? s_ok : s_bad) + edgeR + lastL[hs_len - 2];
      if (v > best) { max_index = hs_len - 1; best = v; }
      for (int i = 1; i < hs_len - 1; i++) {
        if (frow[i] & HIPSTR_ROW_REPEAT) continue;
        const double w = prior + (sx == fseq[i] ? s_ok : s_bad) + lastL[i - 1] + lastR[hs_len - i - 2];
        if (w > best) { max_index = i; best = w; }
      }
    }
    // genomic start() of the oriented blocks
    int32_t start_fw[HIPSTR_MAX_BLOCKS], start_rv[HIPSTR_MAX_BLOCKS];
    const int32_t* bstart = P.block_start + P.locus_block0[pool.locus];
    const int32_t* bend = P.block_ref_end + P.locus_block0[pool.locus];
    for (int b = 0; b < nb; b++) { start_fw[b] = bstart[b]; start_rv[nb - 1 - b] = bend[b] - 1; }
    TAcc acc;
    acc.stutter = P.out_stutter + (size_t)tr * HIPSTR_MAX_BLOCKS_PER_LOCUS;
    acc.lo = P.out_span_start + (size_t)tr * HIPSTR_MAX_BLOCKS_PER_LOCUS;
    acc.hi = P.out_span_len + (size_t)tr * HIPSTR_MAX_BLOCKS_PER_LOCUS;    // holds `hi` until the end
    acc.indels = P.out_indels + (size_t)tr * HIPSTR_MAX_TRACE_INDELS * 2;
    acc.snps = P.out_snps + (size_t)tr * HIPSTR_MAX_TRACE_SNPS * 2;
    acc.n_indels = acc.n_snps = acc.ins = acc.del = 0;
    for (int b = 0; b < HIPSTR_MAX_BLOCKS_PER_LOCUS; b++) { acc.stutter[b] = HIPSTR_NO_STR_DATA; acc.lo[b] = 1 << 30; acc.hi[b] = -1; }
    for (int k = 0; k < HIPSTR_MAX_TRACE_INDELS * 2; k++) acc.indels[k] = 0;
    for (int k = 0; k < HIPSTR_MAX_TRACE_SNPS * 2; k++) acc.snps[k] = 0;
    char* aln = P.out_aln + (size_t)tr * P.aln_stride;
    int n_left = 0;
    // block / offset of a haplotype position
    int fb = 0, fc = max_index;
    while (fc >= P.blocks[hsF.blk_off + fb].len) { fc -= P.blocks[hsF.blk_off + fb].len; fb++; }
    if (max_index == 0) { for (int i = 0; i < seed; i++) aln[n_left++] = 'S'; }
    else {
      const long mi = (long)pitchL * (max_index - 1) + seed - 1;
      if (fc == 0) n_left = t_walk_back(P, L, hsF, start_fw, decL, Ls, Lp, fb - 1, P.blocks[hsF.blk_off + fb - 1].len - 1, mi, aln, acc);
      else n_left = t_walk_back(P, L, hsF, start_fw, decL, Ls, Lp, fb, fc - 1, mi, aln, acc);
      for (int a = 0, z = n_left - 1; a < z; a++, z--) { const char c = aln[a]; aln[a] = aln[z]; aln[z] = c; }   // left side is walked backwards
    }
    if (P.blocks[hsF.blk_off + fb].rep < 0) acc.touch(fb, seed);
    aln[n_left] = 'M';
    const int rmax = hs_len - 1 - max_index;
    int rb = 0, rc = rmax;
    while (rc >= P.blocks[hsR.blk_off + rb].len) { rc -= P.blocks[hsR.blk_off + rb].len; rb++; }
    int n_right = 0;
    char* right = aln + n_left + 1;
    if (rmax == 0) { for (int i = 0; i < n - 1 - seed; i++) right[n_right++] = 'S'; }
    else {
      const long mi = (long)pitchR * (rmax - 1) + (n - 1 - seed) - 1;
      if (rc == 0) n_right = t_walk_back(P, R, hsR, start_rv, decR, Rs, Rp, rb - 1, P.blocks[hsR.blk_off + rb - 1].len - 1, mi, right, acc);
      else n_right = t_walk_back(P, R, hsR, start_rv, decR, Rs, Rp, rb, rc - 1, mi, right, acc);
    }
    right[n_right] = 0;
    P.out_seed_pos[tr] = max_index;
    P.out_flank_ins[tr] = acc.ins; P.out_flank_del[tr] = acc.del;
    P.out_n_indels[tr] = acc.n_indels; P.out_n_snps[tr] = acc.n_snps;
    for (int b = 0; b < HIPSTR_MAX_BLOCKS_PER_LOCUS; b++) {
      const bool any = acc.hi[b] >= acc.lo[b];
      const int lo = acc.lo[b], hi = acc.hi[b];
      acc.lo[b] = any ? lo : 0;            // span_start
      acc.hi[b] = any ? hi - lo + 1 : 0;   // span_len
    }
  }
}

cudaError_t launch_trace(const TraceParams& p, int n_slots, cudaStream_t stream) {
  if (p.n_traces <= 0) return cudaSuccess;
  // 8 resident CTAs of 64 threads per SM (128 registers): more residency at 80 / 64 registers was measured and does
  // not help (profiles/r1_summary.md)
  k_trace<8><<<(n_slots + 63) / 64, 64, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace hipstr
