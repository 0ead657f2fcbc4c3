/*
 * fastapprox.cuh -- device replicas of the single-precision exp/log approximations the
 * reference's log-sum-exp is built on (third-party "fastapprox" by P. Mineiro, vendored by the
 * reference as src/fastonebigheader.h:188-218,320-357; used by mathops.cpp:86-106).
 *
 * They are part of the reference's RESULTS: an exact log-sum-exp differs by up to 5e-2 per call
 * (SURVEY.md 7), so every operation below is an individually rounded binary32 op in the
 * reference's evaluation order -- explicit __f*_rn intrinsics, which the compiler never
 * contracts into FMAs -- with the same conversions (double->float RN, uint->float RN,
 * float->uint/int truncation).
 */
#ifndef HIPSTR_B200_FASTAPPROX_CUH_
#define HIPSTR_B200_FASTAPPROX_CUH_

namespace hipstr {

#define HIPSTR_LOG_THRESH (-6.907755278982137)   /* log(0.001) as glibc rounds it (mathops.h:36) */

// fasterexp = fasterpow2(1.442695040f * p)   (fastonebigheader.h:206-218)
__device__ __forceinline__ float coarse_exp(float p) {
  const float x = __fmul_rn(1.442695040f, p);
  const float c = (x < -126.0f) ? -126.0f : x;
  return __uint_as_float(__float2uint_rz(__fmul_rn(8388608.0f, __fadd_rn(c, 126.94269504f))));
}
// fasterlog   (fastonebigheader.h:348-357)
__device__ __forceinline__ float coarse_log(float x) {
  float y = __uint2float_rn(__float_as_uint(x));
  y = __fmul_rn(y, 8.2629582881927490e-8f);
  return __fsub_rn(y, 87.989971088f);
}
// fastexp = fastpow2(1.442695040f * p)   (fastonebigheader.h:188-204)
__device__ __forceinline__ float fine_exp(float p) {
  const float x = __fmul_rn(1.442695040f, p);
  const float offset = (x < 0.0f) ? 1.0f : 0.0f;
  const float c = (x < -126.0f) ? -126.0f : x;
  const int w = __float2int_rz(c);
  const float z = __fadd_rn(__fsub_rn(c, __int2float_rn(w)), offset);
  float acc = __fadd_rn(c, 121.2740575f);
  acc = __fadd_rn(acc, __fdiv_rn(27.7280233f, __fsub_rn(4.84252568f, z)));
  acc = __fsub_rn(acc, __fmul_rn(1.49012907f, z));
  return __uint_as_float(__float2uint_rz(__fmul_rn(8388608.0f, acc)));
}
// fastlog = 0.69314718f * fastlog2(x)   (fastonebigheader.h:320-337)
__device__ __forceinline__ float fine_log(float x) {
  const unsigned int xi = __float_as_uint(x);
  const float mx = __uint_as_float((xi & 0x007FFFFFu) | 0x3f000000u);
  float y = __uint2float_rn(xi);
  y = __fmul_rn(y, 1.1920928955078125e-7f);
  float l2 = __fsub_rn(y, 124.22551499f);
  l2 = __fsub_rn(l2, __fmul_rn(1.498030302f, mx));
  l2 = __fsub_rn(l2, __fdiv_rn(1.72587999f, __fadd_rn(0.3520887068f, mx)));
  return __fmul_rn(0.69314718f, l2);
}

// fast_log_sum_exp(double, double)   (mathops.cpp:86-95)
__device__ __forceinline__ double lse2(double a, double b) {
  const double hi = a > b ? a : b, lo = a > b ? b : a;
  const double diff = lo - hi;
  if (diff < HIPSTR_LOG_THRESH) return hi;
  return hi + (double)fine_log(__fadd_rn(1.0f, fine_exp(__double2float_rn(diff))));
}

// one term of fast_log_sum_exp(vector)   (mathops.cpp:101-104); the caller adds it to a double
__device__ __forceinline__ double lse_term(double v, double mx) {
  const double diff = v - mx;
  return diff > HIPSTR_LOG_THRESH ? (double)coarse_exp(__double2float_rn(diff)) : 0.0;
}
// the same term when the caller has no use for the clamp of fasterpow2: diff > log(0.001) keeps 1.4427 * diff far
// above -126, so the clamp never fires and the results are identical
__device__ __forceinline__ double lse_term_near(double v, double mx) {
  const double diff = v - mx;
  const float x = __fmul_rn(1.442695040f, __double2float_rn(diff));
  const float e = __uint_as_float(__float2uint_rz(__fmul_rn(8388608.0f, __fadd_rn(x, 126.94269504f))));
  return diff > HIPSTR_LOG_THRESH ? (double)e : 0.0;
}
// ... and when the caller also knows whether the slot holds a term at all: one 32-bit select on the float instead of
// two 64-bit ones ((double)0.0f is +0.0, so the sum sees the same addend)
__device__ __forceinline__ double lse_term_masked(double v, double mx, bool valid) {
  const double diff = v - mx;
  const float x = __fmul_rn(1.442695040f, __double2float_rn(diff));
  const unsigned bits = __float2uint_rz(__fmul_rn(8388608.0f, __fadd_rn(x, 126.94269504f)));
  unsigned kept;
  asm("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %4, 0;\n\tsetp.gt.and.f64 p, %2, %3, q;\n\tselp.b32 %0, %1, 0, p;\n\t}"
      : "=r"(kept) : "r"(bits), "d"(diff), "d"(HIPSTR_LOG_THRESH), "r"((int)valid));
  return (double)__uint_as_float(kept);
}
__device__ __forceinline__ double lse_finish(double mx, double total) {
  return mx + (double)coarse_log(__double2float_rn(total));
}

}  // namespace hipstr
#endif
