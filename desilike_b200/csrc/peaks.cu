// FP64 peak micro-benchmarks: the denominator of the FP64 roofline.  MEASURED_PEAKS.json (driver-written) holds HBM GB/s and
// dense bf16 TFLOP/s only, and the path's arithmetic is FP64, so the library measures what the FP64 pipe of THIS device
// sustains with (kind 0) a register-resident DMMA.8x8x4 loop and (kind 1) a register-resident DFMA loop.
#include "common.h"

namespace dlb {

namespace {

constexpr int PEAK_ITERS = 4096;
constexpr int PEAK_ILP = 8;

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, double seed) {
  double c[PEAK_ILP][2];
  const double a = seed + threadIdx.x * 1e-9, b = 1. - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double seed) {
  double c[PEAK_ILP * 2];
  const double a = seed + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < PEAK_ILP * 2; ++i) c[i] = i;
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP * 2; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < PEAK_ILP * 2; ++i) s += c[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA rate with a limited number of resident warps (one CTA per SM): 28 independent accumulators per warp, as in the
// contraction kernel's main loop.  Tells how many warps per scheduler the FP64 tensor pipe needs to stay busy.
__global__ void dmma_warps_kernel(double* out, double seed) {
  double c[28][2];
  const double a = seed + threadIdx.x * 1e-9, b = 1. - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 28; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 28; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 28; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

int measure_fp64_peak(int kind, double* tflops) {
  int dev = 0, nsm = 0;
  DLB_CUDA(cudaGetDevice(&dev));
  DLB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  if (kind >= 100) {   // kind = 100 + warps per SM
    const int warps = kind - 100;
    if (warps < 1 || warps > 32) { set_error("measure_fp64_peak: 1..32 warps per SM"); return DLB_ERR_ARG; }
    double* out = nullptr;
    DLB_CUDA(cudaMalloc((void**)&out, (size_t)nsm * warps * 32 * sizeof(double)));
    cudaEvent_t e0, e1;
    DLB_CUDA(cudaEventCreate(&e0));
    DLB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      DLB_CUDA(cudaEventRecord(e0));
      dmma_warps_kernel<<<nsm, warps * 32>>>(out, 1.0);
      g_launches++;
      DLB_CUDA(cudaEventRecord(e1));
      DLB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      DLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2 && ms < best) best = ms;
    }
    DLB_CUDA(cudaGetLastError());
    *tflops = (double)PEAK_ITERS * 28 * 512. * nsm * warps / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return DLB_OK;
  }
  const int blocks = nsm * 8, threads = 256;
  double* out = nullptr;
  DLB_CUDA(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  DLB_CUDA(cudaEventCreate(&e0));
  DLB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    DLB_CUDA(cudaEventRecord(e0));
    if (kind == 0) dmma_peak_kernel<<<blocks, threads>>>(out, 1.0);
    else dfma_peak_kernel<<<blocks, threads>>>(out, 1.0);
    g_launches++;
    DLB_CUDA(cudaEventRecord(e1));
    DLB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 2 && ms < best) best = ms;
  }
  DLB_CUDA(cudaGetLastError());
  // flops: DMMA.8x8x4 = 2 * 8*8*4 per warp instruction; DFMA = 2 per thread instruction
  const double per_thread = kind == 0 ? (double)PEAK_ITERS * PEAK_ILP * 512. / 32. : (double)PEAK_ITERS * PEAK_ILP * 2 * 2.;
  *tflops = per_thread * blocks * threads / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return DLB_OK;
}

}  // namespace dlb
