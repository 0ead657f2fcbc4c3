/*
 * nw.cu -- K6: batched Needleman-Wunsch alignment of reads against reference windows, the arithmetic of
 * HipSTR's read left-alignment: NeedlemanWunsch::Align (SeqAlignment/NeedlemanWunsch.cpp:384-423
 * = initMatrices :339-381, nw_helper :193-241, findOptimalStop / findOptimalStopEndPenalty :149-191, traceAlignment
 * :243-337) as called by realign() (SeqAlignment/AlignmentOps.cpp:14-100) and Haplotype::aln_haps_to_ref
 * (SeqAlignment/Haplotype.cpp:58-86).
 *
 * Two phases.  k_nw_fill: one warp per (reference window, read) pair.  The three affine-gap matrices are never
 * materialised: lane k owns a contiguous chunk of reference columns and walks read rows as an anti-diagonal wavefront
 * (at step t it is on row t - k), keeping the previous row of its chunk in shared memory and receiving the cell to its
 * left from lane k - 1 by shuffle.  What the traceback needs -- the three 2-bit predecessor choices of every cell -- is
 * packed into one byte per cell and streamed to global memory.  k_nw_walk: one THREAD per pair walks back along those
 * bytes.  (The first version kept the trace bytes in shared memory and let lane 0 walk back: 36 KB per pair held
 * residency at 4 warps / SM and the walk idled 31 lanes; see profiles/r1_summary.md.)
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
#define NW_RING 40   /* trace rows staged in shared memory: 31 rows of wavefront skew + the 8 flushed together + 1 */

__host__ __device__ inline int nw_pitch(int window_len) { return (window_len + 15) / 16 * 16; }   // bytes per trace row
__host__ __device__ inline size_t nw_ring_offset(int max_ref, int max_read) {
  const size_t rows_and_codes = 3 * sizeof(float) * (size_t)(max_ref + 1) + (size_t)max_ref + (size_t)max_read;
  return (rows_and_codes + 15) / 16 * 16;
}

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

/* Phase 1: the forward pass.  Trace bytes go to global memory ([read row][window column] per pair), so a CTA needs only
 * the rolling rows in shared memory and the SM holds as many warps as the scheduler allows instead of the 4-5 that fit
 * when the trace lives in shared memory.  Lane 0 finally picks the end cell. */
__global__ void __launch_bounds__(32) k_nw_fill(const NwParams P) {
  extern __shared__ unsigned char smem[];
  const int lane = threadIdx.x;
  for (int pair = P.first_pair + blockIdx.x; pair < P.first_pair + P.n_pairs; pair += gridDim.x) {
    const int L1 = P.ref_off[pair + 1] - P.ref_off[pair], L2 = P.read_off[pair + 1] - P.read_off[pair];
    const char* ref = P.ref_seqs + P.ref_off[pair];
    const char* read = P.read_seqs + P.read_off[pair];
    float* rowM = reinterpret_cast<float*>(smem);
    float* rowX = rowM + (P.max_ref + 1);
    float* rowY = rowX + (P.max_ref + 1);
    unsigned char* rcode = reinterpret_cast<unsigned char*>(rowY + (P.max_ref + 1));
    unsigned char* qcode = rcode + P.max_ref;
    // trace bytes are staged in a ring of NW_RING rows in shared memory and leave for global memory as whole rows,
    // 16 bytes per lane per store: byte stores straight from the wavefront would turn every cell into a partial
    // 32-byte sector write in L2 (measured: slower than keeping the whole trace in shared memory)
    const int pitch = nw_pitch(L1);
    unsigned char* ring = smem + nw_ring_offset(P.max_ref, P.max_read);
    unsigned char* trace = P.trace + (P.trace_off[pair] - P.trace_off[P.first_pair]);   // [L2][pitch]
    int flushed = 0;   // rows 1..flushed are in global memory
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
        unsigned char* trow = ring + ((i - 1) % NW_RING) * pitch;
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
      // after step t every lane is done with rows <= t - 30: send them off eight at a time
      if ((t & 7) == 7 || t == steps - 1) {
        __syncwarp();
        const int complete = t == steps - 1 ? L2 : min(L2, t - 30);
        for (int r = flushed + 1; r <= complete; r++) {
          const uint4* src = reinterpret_cast<const uint4*>(ring + ((r - 1) % NW_RING) * pitch);
          uint4* dst = reinterpret_cast<uint4*>(trace + (size_t)(r - 1) * pitch);
          for (int v = lane; v < pitch / 16; v += 32) dst[v] = src[v];
        }
        if (complete > flushed) flushed = complete;
        __syncwarp();
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
      P.end_cell[2 * (size_t)pair] = best_col;
      P.end_cell[2 * (size_t)pair + 1] = kind;
    }
    __syncwarp();
  }
}

/* Phase 2: the walk back (traceAlignment :243-337), one THREAD per pair: a strictly sequential chain of byte loads, so
 * tens of thousands of them in flight hide the memory latency that one lane per warp could not. */
__global__ void __launch_bounds__(128) k_nw_walk(const NwParams P) {
  const int pair = P.first_pair + blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= P.first_pair + P.n_pairs) return;
  const int L1 = P.ref_off[pair + 1] - P.ref_off[pair], L2 = P.read_off[pair + 1] - P.read_off[pair];
  const unsigned char* trace = P.trace + (P.trace_off[pair] - P.trace_off[P.first_pair]);
  const int pitch = nw_pitch(L1);
  char* out = P.out_ops + (size_t)pair * P.ops_stride;
  const int best_col = P.end_cell[2 * (size_t)pair];
  int kind = P.end_cell[2 * (size_t)pair + 1];
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
    const int tb = trace[(size_t)(row - 1) * pitch + (col - 1)];
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

cudaError_t launch_nw(const NwParams& p, int max_ctas, cudaStream_t stream) {
  if (p.n_pairs <= 0) return cudaSuccess;
  const size_t smem = nw_shared_bytes(p.max_ref, p.max_read);
  cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void*>(&k_nw_fill));
  if (e != cudaSuccess) return e;
  const int grid = p.n_pairs < max_ctas ? p.n_pairs : max_ctas;
  k_nw_fill<<<grid, 32, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_nw_walk<<<(p.n_pairs + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

size_t nw_shared_bytes(int max_ref, int max_read) {   // rolling rows of M / X / Y, the two base-code strings, the trace ring
  return nw_ring_offset(max_ref, max_read) + (size_t)NW_RING * nw_pitch(max_ref);
}

}  // namespace hipstr
