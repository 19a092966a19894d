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

// DMMA rate with ILP independent accumulators per warp (dependent chains of length PEAK_ITERS): the latency of the accumulate dependency.
// Tells how few active n-tiles a warp of the contraction kernel can have (masked sub-tiles of triangular blocks) before it becomes
// latency bound: rate(ILP, warps) = min(pipe, ILP * warps_per_scheduler / latency).
template <int ILP>
__global__ void dmma_chain_kernel(double* out, double seed) {
  double c[ILP][2];
  const double a = seed + threadIdx.x * 1e-9, b = 1. - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
int run_dmma_chain(int nsm, int warps, double* tflops) {
  double* out = nullptr;
  DLB_CUDA(cudaMalloc((void**)&out, (size_t)nsm * warps * 32 * sizeof(double)));
  cudaEvent_t e0, e1;
  DLB_CUDA(cudaEventCreate(&e0));
  DLB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    DLB_CUDA(cudaEventRecord(e0));
    dmma_chain_kernel<ILP><<<nsm, warps * 32>>>(out, 1.0);
    g_launches++;
    DLB_CUDA(cudaEventRecord(e1));
    DLB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 2 && ms < best) best = ms;
  }
  DLB_CUDA(cudaGetLastError());
  *tflops = (double)PEAK_ITERS * ILP * 512. * nsm * warps / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return DLB_OK;
}

// DMMA fed from shared memory exactly like the contraction kernel's main loop (warp tile MT x NT m8n8 tiles, 64-bit
// fragment loads from rows of 36 / 76 doubles, optional per-row scaling of the A fragment): the attainable in-loop rate.
//   variant bit 0: scale A fragment (DMUL)   bit 1: skip the shared-memory loads (register operands)   bits 4..: NT (default 7), MT = 2
template <int MT, int NT>
__global__ void dmma_smem_kernel(double* out, int iters, int variant) {
  extern __shared__ double sm[];
  constexpr int AST = 76, BKC = 36;
  double* A_s = sm;                 // [128][AST]
  double* B_s = sm + 128 * AST;     // [16 * NT][BKC]
  for (int i = threadIdx.x; i < 128 * AST + 16 * NT * BKC; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1, lr = lane >> 2, lc = lane & 3;
  const int WMU = (blockDim.x >> 6) * 8;   // rows per unit
  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.;
  const double* A = A_s + (wm * 8 + lr) * AST + lc;
  const double* B = B_s + (wn * NT * 8 + lr) * BKC + lc;
  const bool scale = variant & 1, noload = variant & 2;
  double beta[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) beta[mt] = 1. + 1e-9 * lane * (mt + 1);
  double a[MT], b[NT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) a[mt] = A[(mt * WMU % 128) * AST];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) b[nt] = B[nt * 8 * BKC];
  for (int it = 0; it < iters; ++it) {
#pragma unroll 3
    for (int k4 = 0; k4 < 9; ++k4) {
      if (!noload) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a[mt] = A[(mt * WMU % 128) * AST + k4 * 4 + (it & 1) * 36];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = B[nt * 8 * BKC + k4 * 4];
      }
      if (scale) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a[mt] *= beta[mt];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(a[mt]), "d"(b[nt]));
    }
  }
  double sum = 0.;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) sum += acc[mt][nt][0] + acc[mt][nt][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <int MT, int NT>
int run_dmma_smem(int nsm, int warps, int variant, double* tflops) {
  const int iters = 2000;
  const size_t smem = (size_t)(128 * 76 + 16 * NT * 36) * sizeof(double);
  DLB_CUDA(cudaFuncSetAttribute(dmma_smem_kernel<MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double* out = nullptr;
  DLB_CUDA(cudaMalloc((void**)&out, (size_t)nsm * warps * 32 * sizeof(double)));
  cudaEvent_t e0, e1;
  DLB_CUDA(cudaEventCreate(&e0));
  DLB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    DLB_CUDA(cudaEventRecord(e0));
    dmma_smem_kernel<MT, NT><<<nsm, warps * 32, smem>>>(out, iters, variant);
    g_launches++;
    DLB_CUDA(cudaEventRecord(e1));
    DLB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 2 && ms < best) best = ms;
  }
  DLB_CUDA(cudaGetLastError());
  *tflops = (double)iters * 9 * MT * NT * 512. * nsm * warps / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return DLB_OK;
}

}  // namespace

int measure_fp64_peak(int kind, double* tflops) {
  int dev = 0, nsm = 0;
  DLB_CUDA(cudaGetDevice(&dev));
  DLB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  if (kind >= 2000) {  // kind = 2000 + 100 * ILP + warps per SM: dependent-chain DMMA rate
    const int ilp = (kind - 2000) / 100, warps = kind % 100;
    if (warps < 1 || warps > 32) { set_error("measure_fp64_peak: 1..32 warps per SM"); return DLB_ERR_ARG; }
    switch (ilp) {
      case 1: return run_dmma_chain<1>(nsm, warps, tflops);
      case 2: return run_dmma_chain<2>(nsm, warps, tflops);
      case 4: return run_dmma_chain<4>(nsm, warps, tflops);
      case 6: return run_dmma_chain<6>(nsm, warps, tflops);
      case 10: return run_dmma_chain<10>(nsm, warps, tflops);
      case 14: return run_dmma_chain<14>(nsm, warps, tflops);
      default: set_error("measure_fp64_peak: ILP in {1, 2, 4, 6, 10, 14}"); return DLB_ERR_ARG;
    }
  }
  if (kind >= 1000) {  // kind = 1000 + 100 * tile + 10 * variant-bits + (warps / 4): DMMA fed from shared memory
    const int tile = (kind - 1000) / 100, variant = ((kind - 1000) / 10) % 10, warps = 4 * (kind % 10);
    if (warps < 4 || warps > 32) { set_error("measure_fp64_peak: 4..32 warps"); return DLB_ERR_ARG; }
    if (tile == 0) return run_dmma_smem<2, 7>(nsm, warps, variant, tflops);
    if (tile == 1) return run_dmma_smem<4, 7>(nsm, warps, variant, tflops);
    if (tile == 2) return run_dmma_smem<2, 4>(nsm, warps, variant, tflops);
    if (tile == 3) return run_dmma_smem<1, 7>(nsm, warps, variant, tflops);
    set_error("measure_fp64_peak: unknown tile"); return DLB_ERR_ARG;
  }
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
