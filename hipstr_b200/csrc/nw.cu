/*
 * nw.cu -- K6: batched Needleman-Wunsch alignment of reads against reference windows, the arithmetic of
 * HipSTR's read left-alignment: NeedlemanWunsch::Align (SeqAlignment/NeedlemanWunsch.cpp:384-423
 * = initMatrices :339-381, nw_helper :193-241, findOptimalStop / findOptimalStopEndPenalty :149-191, traceAlignment
 * :243-337) as called by realign() (SeqAlignment/AlignmentOps.cpp:14-100) and Haplotype::aln_haps_to_ref
 * (SeqAlignment/Haplotype.cpp:58-86).
 *
 * One warp per (reference window, read) pair.  The three affine-gap matrices are never materialised: lane k owns a
 * contiguous chunk of reference columns and walks read rows as an anti-diagonal wavefront (at step t it is on row
 * t - k), keeping the previous row of its chunk in shared memory and receiving the cell to its left from lane k - 1 by
 * shuffle.  What the traceback needs -- the three 2-bit predecessor choices of every cell -- is packed into one byte
 * per cell in shared memory (read length x window length bytes, <= 200 KB), so the final, strictly sequential walk
 * back is a chain of shared-memory loads by lane 0, not of DRAM round trips.
 *
 * Scores are the reference's floats (match 2, mismatch -2, gap open 5, gap extend 0.125, "impossible" -1e6): every
 * value that occurs is a multiple of 1/8 below 2^21 in magnitude, hence exact in binary32 in any order, and the
 * reference's own three-way tie rule (bestIndex :125-147) is applied cell by cell -- the operation strings are
 * identical to the reference's, not merely equally good alignments.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace hipstr {

#define NW_MATCH 2.0f
#define NW_MISMATCH (-2.0f)
#define NW_OPEN 5.0f
#define NW_EXTEND 0.125f
#define NW_LARGE 1000000.0f

__device__ __forceinline__ int nw_code(char c) {   // NeedlemanWunsch.cpp:105-123
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

/* bestIndex (:125-147): the second beats the first only if strictly larger, the third beats the second on ties,
 * the first beats the third on ties. */
__device__ __forceinline__ float nw_pick3(float s1, float s2, float s3, int& which) {
  if (s2 > s1) {
    if (s2 > s3) { which = 1; return s2; }
    which = 2;
    return s3;
  }
  if (s3 > s1) { which = 2; return s3; }
  which = 0;
  return s1;
}

__global__ void __launch_bounds__(32) k_nw(const NwParams P) {
  extern __shared__ unsigned char smem[];
  const int lane = threadIdx.x;
  for (int pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
    const int L1 = P.ref_off[pair + 1] - P.ref_off[pair], L2 = P.read_off[pair + 1] - P.read_off[pair];
    const char* ref = P.ref_seqs + P.ref_off[pair];
    const char* read = P.read_seqs + P.read_off[pair];
    char* out = P.out_ops + (size_t)pair * P.ops_stride;
    // shared memory: last computed row of every column (M, X = ref base vs gap, Y = read base vs gap), base codes,
    // then one trace byte per cell
    float* rowM = reinterpret_cast<float*>(smem);
    float* rowX = rowM + (P.max_ref + 1);
    float* rowY = rowX + (P.max_ref + 1);
    unsigned char* rcode = reinterpret_cast<unsigned char*>(rowY + (P.max_ref + 1));
    unsigned char* qcode = rcode + P.max_ref;
    unsigned char* trace = qcode + P.max_read;   // [L2][L1], row-major, rows 1..L2 / columns 1..L1
    for (int j = lane; j < L1; j += 32) rcode[j] = (unsigned char)nw_code(ref[j]);
    for (int i = lane; i < L2; i += 32) qcode[i] = (unsigned char)nw_code(read[i]);
    // row 0 (initMatrices): a leading gap in the read is free unless the reference end is penalised
    for (int j = lane; j <= L1; j += 32) {
      rowM[j] = j == 0 ? 0.0f : -NW_LARGE;
      rowX[j] = j == 0 ? -NW_LARGE : (P.use_ref_end_penalty ? -NW_OPEN - (j - 1) * NW_EXTEND : 0.0f);
      rowY[j] = -NW_LARGE;
    }
    __syncwarp();
    const int chunk = (L1 + 31) / 32;
    const int j0 = 1 + lane * chunk, j1 = min(L1, j0 + chunk - 1);   // this lane's columns [j0, j1] (1-based), may be empty
    // cells handed over by the lane to the left: (row i, column j0-1) arrives each step; the one before is the diagonal
    float diagM = 0.f, diagX = 0.f, diagY = 0.f;
    const int steps = L2 + 31;
    float outM = 0.f, outX = 0.f, outY = 0.f;   // this lane's right-most cell of the row it just finished
    for (int t = 0; t < steps; t++) {
      const int i = t - lane + 1;   // 1-based read row of this lane at this step
      float leftM = __shfl_up_sync(0xffffffffu, outM, 1), leftX = __shfl_up_sync(0xffffffffu, outX, 1),
            leftY = __shfl_up_sync(0xffffffffu, outY, 1);
      if (lane == 0) {              // column 0: only a run of read bases against gaps is possible
        leftM = -NW_LARGE; leftX = -NW_LARGE; leftY = -NW_OPEN - (i - 1) * NW_EXTEND;
      }
      if (i >= 1 && i <= L2) {
        if (i == 1) {               // diagonal of the first row = row 0 of column j0 - 1
          if (j0 - 1 == 0) { diagM = 0.0f; diagX = -NW_LARGE; diagY = -NW_LARGE; }
          else { diagM = -NW_LARGE; diagX = P.use_ref_end_penalty ? -NW_OPEN - (j0 - 2) * NW_EXTEND : 0.0f; diagY = -NW_LARGE; }
        }
        float dM = diagM, dX = diagX, dY = diagY, lM = leftM, lX = leftX, lY = leftY;
        const int q = qcode[i - 1];
        unsigned char* trow = trace + (size_t)(i - 1) * L1;
        for (int j = j0; j <= j1; j++) {
          const float uM = rowM[j], uX = rowX[j], uY = rowY[j];   // row i-1 of this column
          const int r = rcode[j - 1];
          int c0, c1, c2;
          const float m = nw_pick3(dM, dX, dY, c0) + ((r == 4 || q == 4 || r == q) ? NW_MATCH : NW_MISMATCH);
          const float x = nw_pick3(lM - NW_OPEN, lX - NW_EXTEND, lY - NW_OPEN, c1);
          const float y = nw_pick3(uM - NW_OPEN, uX - NW_OPEN, uY - NW_EXTEND, c2);
          trow[j - 1] = (unsigned char)(c0 | (c1 << 2) | (c2 << 4));
          rowM[j] = m; rowX[j] = x; rowY[j] = y;
          dM = uM; dX = uX; dY = uY;
          lM = m; lX = x; lY = y;
        }
        diagM = leftM; diagX = leftX; diagY = leftY;   // (row i, column j0-1) is the diagonal of the next row
        if (j1 >= j0) { outM = lM; outX = lX; outY = lY; }
        else { outM = leftM; outX = leftX; outY = leftY; }   // an empty chunk passes its input through
      }
    }
    __syncwarp();
    if (lane == 0) {
      // where the alignment ends (last read row): the corner, or the best cell of the row (findOptimalStop :149-173)
      int best_col = L1, kind = 0;
      float best;
      if (P.use_ref_end_penalty) {
        best = rowM[L1];
        if (rowX[L1] > best) { best = rowX[L1]; kind = 1; }
        if (rowY[L1] > best) { best = rowY[L1]; kind = 2; }
      } else {
        best = -NW_LARGE; best_col = -1; kind = -1;
        for (int j = 0; j <= L1; j++) {
          const float m = j == 0 ? -NW_LARGE : rowM[j], x = j == 0 ? -NW_LARGE : rowX[j],
                      y = j == 0 ? -NW_OPEN - (L2 - 1) * NW_EXTEND : rowY[j];
          if (m >= best) { best = m; best_col = j; kind = 0; }
          if (x > best) { best = x; best_col = j; kind = 1; }
          if (y > best) { best = y; best_col = j; kind = 2; }
        }
      }
      P.out_score[pair] = best;
      // walk back (traceAlignment :243-337), emitting operations last to first
      int n = 0, row = L2, col = best_col;
      bool ok = true;
      for (int j = L1; j > best_col; j--) out[n++] = 'D';   // trailing reference bases
      while (row > 0) {
        // column 0 holds only the leading run of read bases (trace 2); its M / X cells are impossible
        if (col == 0) {
          if (kind != 2) { ok = false; break; }
          out[n++] = 'I'; row--;
          continue;
        }
        const int tb = trace[(size_t)(row - 1) * L1 + (col - 1)];
        if (kind == 0) { out[n++] = 'M'; kind = tb & 3; row--; col--; }
        else if (kind == 1) { out[n++] = 'D'; kind = (tb >> 2) & 3; col--; }
        else if (kind == 2) { out[n++] = 'I'; kind = (tb >> 4) & 3; row--; }
        else { ok = false; break; }
      }
      for (; col > 0; col--) out[n++] = 'D';                // leading reference bases
      for (int a = 0, z = n - 1; a < z; a++, z--) { const char c = out[a]; out[a] = out[z]; out[z] = c; }
      out[n] = 0;
      P.out_len[pair] = ok ? n : -1;
    }
    __syncwarp();
  }
}

cudaError_t launch_nw(const NwParams& p, int max_ctas, cudaStream_t stream) {
  if (p.n_pairs <= 0) return cudaSuccess;
  const size_t smem = 3 * sizeof(float) * (size_t)(p.max_ref + 1) + (size_t)p.max_ref + (size_t)p.max_read + (size_t)p.max_ref * p.max_read;
  cudaError_t e = cudaFuncSetAttribute(k_nw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = p.n_pairs < max_ctas ? p.n_pairs : max_ctas;
  k_nw<<<grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

size_t nw_shared_bytes(int max_ref, int max_read) {
  return 3 * sizeof(float) * (size_t)(max_ref + 1) + (size_t)max_ref + (size_t)max_read + (size_t)max_ref * max_read;
}

}  // namespace hipstr
