/*
 * snp_phase.cu -- K7: phasing log-likelihoods of reads from the phased heterozygous SNPs they overlap.
 *
 * Replaces calc_het_snp_factors / add_log_phasing_probs / extract_bases_and_qualities
 * (src/snp_phasing_quality.cpp:4-120) and SNPTree::findContained (src/snp_tree.h:114-126) for every read
 * of a batch of loci at once.  The per-sample interval tree of the reference becomes one position-sorted
 * SNP array per sample ("SNP set"): findContained(start, stop) returns exactly the SNPs with
 * start <= pos <= stop in position order, i.e. a lower bound followed by a forward scan.
 *
 * One thread per ENTRY (an STR read, optionally followed by its mate): it walks the alignments of the
 * entry in order, and for each one the CIGAR and the SNP list side by side, adding the quality terms to
 * the two running doubles in the reference's order -- the sums are therefore bit-identical
 * (IEEE double additions of host-computed table entries).  Work per entry is a handful of bytes
 * (CIGAR, the bases under the SNPs) so the kernel is bound by its scattered loads, not by arithmetic;
 * the grid is sized to the SM count and strides over the entries.
 */
#include "kernels.h"

namespace hipstr {

namespace {

__device__ __forceinline__ int lower_bound_pos(const uint32_t* pos, int lo, int hi, uint32_t key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(pos + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) k_snp_phase(const SnpPhaseParams P) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_entries; e += gridDim.x * blockDim.x) {
    double log_p1 = 0.0, log_p2 = 0.0;
    int n1 = 0, n2 = 0, n_mis = 0, status = 0;
    const int set = P.entry_snp_set[e];
    if (set >= 0) {
      const int set_lo = P.set_off[set], set_hi = P.set_off[set + 1];
      for (int a = P.entry_aln_off[e]; a < P.entry_aln_off[e + 1]; a++) {
        const int32_t aln_pos = P.aln_pos[a];
        // GetEndPosition() is exclusive: only SNPs the read overlaps (snp_phasing_quality.cpp:66-67)
        const uint32_t last = (uint32_t)(P.aln_end[a] - 1);
        int s = lower_bound_pos(P.snp_pos, set_lo, set_hi, (uint32_t)aln_pos);
        if (s == set_hi || __ldg(P.snp_pos + s) > last) continue;
        const char* bases = P.bases + P.aln_seq_off[a];
        const unsigned char* quals = (const unsigned char*)P.quals + P.aln_seq_off[a];
        const int n_bases = P.aln_seq_off[a + 1] - P.aln_seq_off[a];
        int c = P.aln_cigar_off[a];
        const int c_end = P.aln_cigar_off[a + 1];
        int32_t pos = aln_pos;
        uint32_t base_index = 0;
        uint32_t snp = __ldg(P.snp_pos + s);
        // extract_bases_and_qualities (:4-63), consuming each SNP as soon as it is resolved
        while (c < c_end) {
          const char type = P.cigar_type[c];
          const int32_t len = P.cigar_len[c];
          int64_t idx = -2;   // -2: op consumed without resolving the SNP; -1: SNP resolved as "no base"
          if (type == 'M' || type == '=' || type == 'X') {
            if (snp < (uint32_t)(pos + len)) idx = (int64_t)(snp - (uint32_t)pos + base_index);
            else { pos += len; base_index += len; c++; }
          } else if (type == 'D') {
            if (snp < (uint32_t)(pos + len)) idx = -1;
            else { pos += len; c++; }
          } else if (type == 'I') { base_index += len; c++; }
          else if (type == 'S') {
            if (snp < (uint32_t)pos) idx = -1;   // soft-clipped bases are ignored
            else { base_index += len; c++; }
          } else if (type == 'H') c++;
          else { status = 1; break; }            // the reference dies on any other CIGAR character
          if (idx == -2) continue;
          if (idx >= 0) {
            if (idx >= n_bases) { status = 2; break; }   // std::string::at would throw
            const char b = bases[idx];
            if (b != '-') {
              const double ok = __ldg(P.qual_lut + 2 * quals[idx]), bad = __ldg(P.qual_lut + 2 * quals[idx] + 1);
              if (b == P.snp_base1[s]) { log_p1 += ok; log_p2 += bad; n1++; }
              else if (b == P.snp_base2[s]) { log_p1 += bad; log_p2 += ok; n2++; }
              else { log_p1 += bad; log_p2 += bad; n_mis++; }
            }
          }
          s++;
          if (s == set_hi) break;
          snp = __ldg(P.snp_pos + s);
          if (snp > last) break;
        }
        // the reference asserts that every overlapped SNP was resolved before the CIGAR ran out
        if (status == 0 && c >= c_end && s < set_hi && __ldg(P.snp_pos + s) <= last) status = 3;
        if (status) break;
      }
    }
    P.out_log_p1[e] = log_p1;
    P.out_log_p2[e] = log_p2;
    P.out_counts[4 * e + 0] = n1;
    P.out_counts[4 * e + 1] = n2;
    P.out_counts[4 * e + 2] = n_mis;
    P.out_counts[4 * e + 3] = status;
  }
}

}  // namespace

cudaError_t launch_snp_phase(const SnpPhaseParams& p, int n_sm, cudaStream_t stream) {
  if (p.n_entries <= 0) return cudaSuccess;
  const int threads = 128;
  const int needed = (p.n_entries + threads - 1) / threads;
  const int grid = needed < n_sm * 16 ? needed : n_sm * 16;
  k_snp_phase<<<grid, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace hipstr
