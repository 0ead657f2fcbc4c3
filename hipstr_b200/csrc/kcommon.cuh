/* kcommon.cuh -- small PTX helpers shared by the alignment kernels (K1a stutter.cu, K1b kernels.cu). */
#ifndef HIPSTR_B200_KCOMMON_CUH_
#define HIPSTR_B200_KCOMMON_CUH_
#include <cuda_runtime.h>
#include <stdint.h>

namespace hipstr {

#define FULL 0xffffffffu
#define IMPOSSIBLE (-1000000000.0)            /* HapAligner.cpp:20 */

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// Shared-memory accesses of the hot loops use explicit 32-bit shared-window addresses: the compiler
// otherwise rebuilds the generic->shared base (S2UR/ULEA) inside the loop.
// a >= 0 ? x : y as one select the compiler cannot turn back into a branch around the code that computes x (it otherwise
// sinks the loads feeding x into that branch, where each waits for the previous one's consumer)
__device__ __forceinline__ double select_if_nonneg(int a, double x, double y) {
  double v;
  asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\tselp.f64 %0, %2, %3, p;\n\t}" : "=d"(v) : "r"(a), "d"(x), "d"(y));
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v));
}

#define HIPSTR_COL_BYTES (HIPSTR_VAL_STRIDE * 8)

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a pooled read's packed bases + qualities ------
// One elected lane arms an mbarrier with the byte count and issues the two bulk copies; the warp
// then waits on the barrier's phase.  Sources are 16-byte aligned and padded by the host lowering.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// The predecessor choices of HapAligner::retrace (HapAligner.cpp:345-361): within TRACE_LL_TOL = 0.001 the left side of
// the seed prefers the later candidate, the right side (aligned reversed) the earlier one.
#define HIPSTR_TRACE_TOL 0.001
__device__ __forceinline__ int pick3(bool rev, double v1, double v2, double v3) {
  if (!rev) {
    if (v1 > v2 + HIPSTR_TRACE_TOL) return v1 > v3 + HIPSTR_TRACE_TOL ? 0 : 2;
    return v2 > v3 + HIPSTR_TRACE_TOL ? 1 : 2;
  }
  if (v3 > v2 + HIPSTR_TRACE_TOL) return v3 > v1 + HIPSTR_TRACE_TOL ? 2 : 0;
  return v2 > v1 + HIPSTR_TRACE_TOL ? 1 : 0;
}
__device__ __forceinline__ int pick2(bool rev, double v1, double v2) {
  if (!rev) return v1 > v2 + HIPSTR_TRACE_TOL ? 0 : 1;
  return v2 > v1 + HIPSTR_TRACE_TOL ? 1 : 0;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace hipstr
#endif
