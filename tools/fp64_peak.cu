/*
 * fp64_peak.cu -- FP64 issue-rate microbenchmark for the roofline of K1 ("the honest compute ceiling must
 * be measured on the box").  K1's flank recurrence is DADD + double max (no FMA: the reference's operation order is kept),
 * so the ceilings that matter are the chip-wide rates of
 *     dadd     independent double additions            (DADD)
 *     dmax     independent double maxima                (DSETP + selects, or DMNMX where the ISA has it)
 *     cell     the flank cell of HapAligner.cpp:141-153: 9 adds + 4 maxima per cell, as K1 evaluates it
 * each with 8 independent chains per thread and every SM full (2048 threads), and -- because K1 is latency bound --
 *     dadd_chain  ONE dependent DADD chain per thread (latency x residency instead of issue rate).
 * Prints one JSON object; bench.py reads the committed copy (profiles/fp64_peak.json) for roofline.peak.
 *
 * Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/fp64_peak tools/fp64_peak.cu
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 4096

__global__ void k_dadd(double* out, double a) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 1e-9 + k;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = __dadd_rn(x[k], a);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

__global__ void k_dmax(double* out, const double* in) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 1e-9 + k;
  double b0 = in[0], b1 = in[1];
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = dmax(x[k], (i & 1) ? b0 : b1);
    b0 += 1.0; b1 += 1.0;   // keeps the maxima from being hoisted; 2 adds per 8 maxima are counted below
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

/* 4 independent flank cells per iteration, each 9 DADD + 4 max like k_align's inner loop */
__global__ void k_cell(double* out, const double* in) {
  double M[4], I[4], D[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { M[k] = -1.0 - k - threadIdx.x * 1e-6; I[k] = -2.0 - k; D[k] = -3.0 - k; }
  const double m2m = in[0], m2i = in[1], m2d = in[2], i2m = in[3], i2i = in[4], d2m = in[5], d2d = in[6], e = in[7];
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double Mn = e + dmax(I[k] + m2i, dmax(M[k] + m2m, D[k] + m2d));
      const double In = e + dmax(M[k] + i2m, I[k] + i2i);
      const double Dn = dmax(M[k] + d2m, D[k] + d2d);
      M[k] = Mn; I[k] = In; D[k] = Dn;
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += M[k] + I[k] + D[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dadd_chain(double* out, double a) {
  double x = threadIdx.x * 1e-9;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) x = __dadd_rn(x, a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <class F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; w++) launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    fprintf(stderr, "no CUDA device\n");
    return 1;
  }
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int threads = 256, blocks = sms * 8;   // 2048 threads per SM
  const double n_threads = (double)threads * blocks;
  double *out, *in;
  cudaMalloc(&out, sizeof(double) * threads * blocks);
  cudaMalloc(&in, 64);
  const double h_in[8] = {-0.01, -4.0, -4.0, -0.4586751453870818910216436, -1.0, -0.4586751453870818910216436, -1.0, -0.001};
  cudaMemcpy(in, h_in, 64, cudaMemcpyHostToDevice);

  const double ms_add = time_ms([&] { k_dadd<<<blocks, threads>>>(out, 1e-3); });
  const double ms_max = time_ms([&] { k_dmax<<<blocks, threads>>>(out, in); });
  const double ms_cell = time_ms([&] { k_cell<<<blocks, threads>>>(out, in); });
  const double ms_chain = time_ms([&] { k_dadd_chain<<<blocks, threads>>>(out, 1e-3); });
  // the chain at K1's residency: 16 warps per SM
  const double ms_chain16 = time_ms([&] { k_dadd_chain<<<sms * 2, 256>>>(out, 1e-3); });
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }

  const double add_ops = n_threads * ITERS * 8.0;
  const double max_ops = n_threads * ITERS * 8.0;
  const double cells = n_threads * ITERS * 4.0;
  const double chain16_ops = (double)sms * 2 * 256 * ITERS * 8.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_attr\": %.0f,\n"
         " \"dadd_per_s\": %.4e, \"dmax_per_s\": %.4e, \"flank_cells_per_s\": %.4e, \"flank_cell_fp64_ops_per_s\": %.4e,\n"
         " \"dadd_chain_per_s_2048_threads_per_sm\": %.4e, \"dadd_chain_per_s_16_warps_per_sm\": %.4e,\n"
         " \"dadd_per_clk_per_sm\": %.2f, \"dmax_per_clk_per_sm\": %.2f, \"flank_cells_per_clk_per_sm\": %.3f,\n"
         " \"dadd_chain_latency_clk_at_16_warps\": %.2f,\n"
         " \"how\": \"8 independent chains per thread, 2048 threads per SM, %d iterations, best of 5 (CUDA events); per-clock figures use the attribute clock\"}\n",
         prop.name, sms, khz / 1e3, add_ops / (ms_add * 1e-3), max_ops / (ms_max * 1e-3), cells / (ms_cell * 1e-3),
         cells * 13.0 / (ms_cell * 1e-3), add_ops / (ms_chain * 1e-3), chain16_ops / (ms_chain16 * 1e-3),
         add_ops / (ms_add * 1e-3) / (khz * 1e3) / sms, max_ops / (ms_max * 1e-3) / (khz * 1e3) / sms,
         cells / (ms_cell * 1e-3) / (khz * 1e3) / sms,
         /* one warp issues 32 chain adds per latency; 16 warps share 4 schedulers */
         (ms_chain16 * 1e-3) * (khz * 1e3) / (ITERS * 8.0), ITERS);
  cudaFree(out); cudaFree(in);
  return 0;
}
