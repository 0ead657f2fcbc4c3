/*
 * trace.cu -- K5, second half: the walk back of HapAligner::retrace (SeqAlignment/HapAligner.cpp:363-571).
 *
 * The forward pass of a trace is K1a + K1b in TRACE mode (stutter.cu, kernels.cu): one warp per trace runs the same
 * wavefront DP as the alignment kernel and leaves, per trace,
 *   - one predecessor-choice byte per flank cell (bits 0-1: match state came from insertion-left / deletion-diagonal /
 *     match-diagonal; bit 2: deletion state came from a match; bit 3: insertion state came from a match) -- the reference
 *     re-derives these from its three full matrices while walking; they only depend on values at hand in the forward pass,
 *   - the best artifact size and position of every repeat-block column (HapAligner.cpp:79-97),
 *   - the haplotype position of the seed base (compute_aln_logprob, :163-231).
 * k_trace_walk: ONE THREAD per trace follows the bytes from the seed outwards on both sides -- a strictly sequential
 * chain of byte loads, tens of thousands of them in flight -- and emits what AlignmentTrace holds
 * (SeqAlignment/AlignmentTraceback.h:10-115): the read-vs-haplotype operations, stutter sizes, the read span of every
 * block, flank indels and flank SNPs.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/hipstr_b200.h"
#include "kernels.h"
#include "layout.h"

namespace hipstr {

#define T_MIN_SNP_LOG_CORRECT (-0.0043648054) /* HapAligner.cpp:24 */

// One side of the seed: column k of the side is read base (rev ? n_read-1-k : k).
struct TSide {
  const uint8_t* bases;   // codes, read order
  const uint8_t* quals;
  const double* qual_lut;
  int n, rev, n_read;
  int pitch;              // row pitch of this side's decision matrix = n
  __device__ __forceinline__ int ridx(int k) const { return rev ? n_read - 1 - k : k; }
  __device__ __forceinline__ int code(int k) const { return bases[ridx(k)]; }
  __device__ __forceinline__ double lc(int k) const { return __ldg(qual_lut + 2 * quals[ridx(k)]); }
};
typedef const unsigned char* BMat;
typedef const int32_t* IMat;
typedef TraceWalkParams TraceParams;

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
                           BMat dec, IMat art_size, IMat art_pos,
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

__global__ void __launch_bounds__(128) k_trace_walk(const TraceWalkParams P) {
  const int tr = blockIdx.x * blockDim.x + threadIdx.x;
  if (tr >= P.n_traces) return;
  const DevPool pool = P.pools[P.trace_pool[tr]];
  const int h = P.trace_hap[tr];
  const DevHapSide hsF = P.hapsides[pool.hap_rec0 + 2 * h];
  const DevHapSide hsR = P.hapsides[pool.hap_rec0 + 2 * h + 1];
  const int n = pool.len, seed = pool.seed, nL = seed, nR = n - seed - 1, hs_len = hsF.len, nb = hsF.n_blocks;
  const unsigned char* decL = P.dec + P.dec_off[tr];
  const unsigned char* decR = decL + (size_t)hs_len * nL;
  const int32_t* Ls = P.art + P.art_off[tr];            // sizes [left: blocks x nL][right: blocks x nR], then positions
  const int32_t* Rs = Ls + (size_t)nb * nL;
  const int32_t* Lp = Ls + (size_t)nb * (nL + nR);
  const int32_t* Rp = Lp + (size_t)nb * nL;
  TSide L, R;
  L.bases = (const uint8_t*)P.bases + pool.seq_off;
  L.quals = (const uint8_t*)P.quals + pool.seq_off;
  L.qual_lut = P.qual_lut;
  L.n = nL; L.rev = 0; L.n_read = n; L.pitch = nL;
  R = L; R.n = nR; R.rev = 1; R.pitch = nR;
  const int max_index = P.seed_pos[tr];
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
    const long mi = (long)nL * (max_index - 1) + seed - 1;
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
    const long mi = (long)nR * (rmax - 1) + (n - 1 - seed) - 1;
    if (rc == 0) n_right = t_walk_back(P, R, hsR, start_rv, decR, Rs, Rp, rb - 1, P.blocks[hsR.blk_off + rb - 1].len - 1, mi, right, acc);
    else n_right = t_walk_back(P, R, hsR, start_rv, decR, Rs, Rp, rb, rc - 1, mi, right, acc);
  }
  right[n_right] = 0;
  P.out_seed_pos[tr] = max_index;
  P.out_flank_ins[tr] = acc.ins; P.out_flank_del[tr] = acc.del;
  P.out_n_indels[tr] = acc.n_indels; P.out_n_snps[tr] = acc.n_snps;   // the true counts; only the first MAX entries are stored
  for (int b = 0; b < HIPSTR_MAX_BLOCKS_PER_LOCUS; b++) {
    const bool any = acc.hi[b] >= acc.lo[b];
    const int lo = acc.lo[b], hi = acc.hi[b];
    acc.lo[b] = any ? lo : 0;            // span_start
    acc.hi[b] = any ? hi - lo + 1 : 0;   // span_len
  }
}

cudaError_t launch_trace_walk(const TraceWalkParams& p, cudaStream_t stream) {
  if (p.n_traces <= 0) return cudaSuccess;
  k_trace_walk<<<(p.n_traces + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace hipstr
