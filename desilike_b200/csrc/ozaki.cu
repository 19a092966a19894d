// FP64-grade GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM accumulators): Ozaki-style exact integer slicing.
//
// Where it is used: the hidden layers of WIDE MLP emulators (reference emulators/conversion.py:18-33: v <- act(v @ kernel + bias)), which the
// FP64 DMMA pipe runs at ~18 TFLOP/s; opt-in (dlb_plan_desc::mlp_slices), parity-gated by |delta chi2| < 1e-4 (BASELINE.json north_star allows
// reduced-precision tensor-core arithmetic under that tolerance; 6 slices reproduce FP64 to ~1e-11 relative, see tests/test_gpu_ozaki.py).
//
// Scheme.  Every row a of A (and column b of B) is scaled by a power of two, |a| 2^-Ea <= 1/2, and cut into `ns` signed 7-bit slices
//     a 2^-Ea = sum_s qa_s 2^(-7 (s + 1)) + tail,   qa_s = rint(.) in [-64, 64]   (exact in FP64: every step is a power-of-two scaling and a subtraction)
// so that  A . B = 2^(Ea + Eb) sum_t 2^(-7 (t + 2)) sum_{i + j = t} (qa_i . qb_j)  with the slice products for i + j <= ns - 1.  Each int8 GEMM
// accumulates EXACTLY in the int32 TMEM accumulators (K (t + 1) 64^2 < 2^31), one accumulator per t; the FP64 epilogue adds the `ns` accumulators
// smallest first, applies the exponents, bias and activation.  The only error is the dropped tail ~ K 2^(-7 (ns + 1)).
//
// Kernels
//   ozaki_slice_kernel : FP64 rows -> int8 slices written in the tensor core's canonical K-major no-swizzle shared-memory image
//                        (8 rows x 16 bytes core matrices), tiled [row tile of 128][k block of 64][slice] so that one stage is ONE bulk copy.
//   ozaki_gemm_kernel  : persistent, warp-specialised: warp 4 = TMA producer (cp.async.bulk + mbarrier ring), warp 5 = MMA issuer (one elected
//                        thread, tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 64, K = 32 per instruction, ns accumulators of 64 TMEM columns),
//                        warps 0-3 = epilogue (tcgen05.ld 32x32b -> FP64 combine -> bias / activation -> global).
#include "common.h"

#include <cstring>

namespace dlb {

namespace {

constexpr int OZ_M = 128, OZ_N = 64, OZ_KB = 32;          // output tile, k-block (bytes = int8 elements) per pipeline stage = one MMA K
constexpr int OZ_A_SLICE = OZ_M * OZ_KB, OZ_B_SLICE = OZ_N * OZ_KB;   // bytes of one slice of one stage
constexpr int OZ_MAXS = 8;
constexpr int OZ_THREADS = 192;

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// bounded wait: a pipeline bug must not hang the GPU (the box would be lost); trap instead -> the launch fails with an error
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
  for (unsigned it = 0; it < (1u << 28); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// canonical K-major, no-swizzle operand image: [k chunk of 16 bytes][8-row group][row in group][16 bytes]
__host__ __device__ __forceinline__ size_t oz_tile_offset(int rows_in_tile, int r, int k) {
  return (size_t)(k >> 4) * (rows_in_tile * 16) + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16 + (k & 15);
}

// ---- slicing: one warp per row, lane l owns the 16-byte k chunks l, l + 32, ... ----
__global__ void __launch_bounds__(128) ozaki_slice_kernel(const double* __restrict__ X, long long rows, long long rows_pad, int K, int Kpad, long long ld, int ns,
                                                          int8_t* __restrict__ slices, int* __restrict__ exps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + warp;
  if (row >= rows_pad) return;
  const int nkb = Kpad / OZ_KB;
  const long long rt = row / OZ_M;
  const int rl = (int)(row - rt * OZ_M);
  const double* x = X + row * ld;
  double amax = 0.;
  if (row < rows)
    for (int k = lane; k < K; k += 32) amax = fmax(amax, fabs(x[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  int E = 0;
  if (amax > 0. && isfinite(amax)) { int e; frexp(amax, &e); E = e + 1; }   // amax = m 2^e, m in [0.5, 1)  ->  amax 2^-(e + 1) < 1/2
  if (lane == 0) exps[row] = E;
  for (int c = lane; c < Kpad / 16; c += 32) {
    const int kb = c / (OZ_KB / 16), kc = c % (OZ_KB / 16);
    uint32_t q[OZ_MAXS][4];
#pragma unroll
    for (int s = 0; s < OZ_MAXS; ++s) q[s][0] = q[s][1] = q[s][2] = q[s][3] = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = c * 16 + j;
      double r = (row < rows && k < K) ? scalbn(x[k], -E) : 0.;
      if (!isfinite(r)) r = 0.;
#pragma unroll
      for (int s = 0; s < OZ_MAXS; ++s)
        if (s < ns) {
          r *= 128.;
          const double qs = rint(r);
          r -= qs;
          q[s][j >> 2] |= ((uint32_t)(int)qs & 0xFFu) << (8 * (j & 3));
        }
    }
    int8_t* base = slices + ((size_t)(rt * nkb + kb) * ns) * OZ_A_SLICE + (size_t)kc * (OZ_M * 16) + (size_t)(rl >> 3) * 128 + (size_t)(rl & 7) * 16;
#pragma unroll
    for (int s = 0; s < OZ_MAXS; ++s)
      if (s < ns) *reinterpret_cast<uint4*>(base + (size_t)s * OZ_A_SLICE) = make_uint4(q[s][0], q[s][1], q[s][2], q[s][3]);
  }
}

struct OzakiParams {
  const int8_t* A;        // [row tiles][nkb][ns][OZ_A_SLICE]
  const int8_t* B;        // [col tiles][nkb][ns][OZ_B_SLICE]   (plan constant)
  const int* Ea;          // [rows_pad]
  const double* Sb;       // [ncols_pad] 2^Eb as doubles (epilogue)
  const double* bias;     // [ncols] or NULL
  double* out; long long ldo; long long rows; int ncols;
  int nkb, ns, act, ntm, ntn;
  int nstages;
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE; LBO = byte distance between the two 16-byte k chunks of one
// MMA, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ double oz_silu(double v) { return v * fast_rcp(1. + fast_exp(fmin(fmax(-v, -708.), 708.))); }

__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ OzakiParams p) {
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  const int ns = p.ns, S = p.nstages;
  const uint32_t stage_bytes = (uint32_t)ns * (OZ_A_SLICE + OZ_B_SLICE);
  unsigned char* stages = oz_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + (size_t)S * stage_bytes);   // full[S], empty[S], tmem_full, tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = oz_smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t tmem_full = bar0 + 8u * (2 * S), tmem_empty = bar0 + 8u * (2 * S + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full(s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty(s)));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tmem_full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(tmem_empty));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (one CTA per SM); the allocating warp also frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(oz_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const long long ntiles = (long long)p.ntm * p.ntn;
  if (warp == 4) {
    // ===== producer: one bulk copy of all A slices + one of all B slices per (tile, k block) =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long tm = tile / p.ntn; const int tn = (int)(tile - tm * p.ntn);
        for (int kb = 0; kb < p.nkb; ++kb) {
          oz_mbar_wait(empty(s), ph ^ 1u);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full(s)), "r"(stage_bytes) : "memory");
          const int8_t* srcA = p.A + ((size_t)(tm * p.nkb + kb) * ns) * OZ_A_SLICE;
          const int8_t* srcB = p.B + ((size_t)((long long)tn * p.nkb + kb) * ns) * OZ_B_SLICE;
          const uint32_t dst = oz_smem_u32(stages + (size_t)s * stage_bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(srcA),
                       "r"((uint32_t)ns * OZ_A_SLICE), "r"(full(s)) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (uint32_t)ns * OZ_A_SLICE), "l"(srcB),
                       "r"((uint32_t)ns * OZ_B_SLICE), "r"(full(s)) : "memory");
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer: one elected thread =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
      int s = 0; uint32_t ph = 0, tph = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        oz_mbar_wait(tmem_empty, tph ^ 1u);   // the epilogue has drained the accumulators of the previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < p.nkb; ++kb) {
          oz_mbar_wait(full(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = oz_smem_u32(stages + (size_t)s * stage_bytes), b0 = a0 + (uint32_t)ns * OZ_A_SLICE;
#pragma unroll 1
          for (int k32 = 0; k32 < OZ_KB / 32; ++k32) {
            for (int i = 0; i < ns; ++i) {
              const uint64_t da = oz_desc(a0 + (uint32_t)i * OZ_A_SLICE + (uint32_t)k32 * 2 * (OZ_M * 16), OZ_M * 16, 128);
              for (int j = 0; i + j < ns; ++j) {
                const uint64_t db = oz_desc(b0 + (uint32_t)j * OZ_B_SLICE + (uint32_t)k32 * 2 * (OZ_N * 16), OZ_N * 16, 128);
                const uint32_t d = tmem + (uint32_t)(i + j) * OZ_N;
                // accumulate unless this is the first product of its accumulator in this tile (i == 0 visits every t = j first)
                const uint32_t acc = (kb > 0 || k32 > 0 || i > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                    ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
              }
            }
          }
          // the stage may be refilled once these MMAs have read it
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty(s)) : "memory");
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tmem_full) : "memory");
        tph ^= 1u;
      }
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (= output rows of the tile) =====
    uint32_t tph = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long tm = tile / p.ntn; const int tn = (int)(tile - tm * p.ntn);
      oz_mbar_wait(tmem_full, tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long row = tm * OZ_M + warp * 32 + lane;
      // ---- phase A (TMEM busy): accumulators -> FP64 in registers.  The int32 accumulators of three consecutive t are first combined EXACTLY in
      //      int64 (G_g = acc_3g 2^14 + acc_3g+1 2^7 + acc_3g+2 < 2^46), then v = sum_g G_g 2^(-7 (3 g + 4)), smallest group first ----
      double v[OZ_N];
      const int ng = (ns + 2) / 3;
#pragma unroll
      for (int c0 = 0; c0 < OZ_N; c0 += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[c0 + j] = 0.;
        for (int g = ng - 1; g >= 0; --g) {
          long long G[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) G[j] = 0;
          for (int u = 0; u < 3; ++u) {
            const int t = 3 * g + u;
            if (t >= ns) break;
            uint32_t r[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * OZ_N + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) G[j] += (long long)(int)r[j] * (1LL << (7 * (2 - u)));
          }
          const double w = scalbn(1., -7 * (3 * g + 4));
#pragma unroll
          for (int j = 0; j < 8; ++j) v[c0 + j] = fma((double)G[j], w, v[c0 + j]);
        }
      }
      // TMEM is free again: the MMA warp may start the next tile while this warp finishes its rows
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty) : "memory");
      // ---- phase B: exponents, bias, activation, store ----
      if (row < p.rows) {
        const double sa = scalbn(1., p.Ea[row]);
        double* orow = p.out + row * p.ldo + (long long)tn * OZ_N;
#pragma unroll
        for (int j = 0; j < OZ_N; ++j) {
          const int col = tn * OZ_N + j;
          if (col < p.ncols) {
            double y = v[j] * sa * p.Sb[col] + (p.bias ? p.bias[col] : 0.);
            if (p.act == 1) y = oz_silu(y);
            else if (p.act == 2) y = fmax(y, 0.);
            else if (p.act == 3) y = tanh(y);
            orow[j] = y;
          }
        }
      }
      tph ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// narrow layers (the first layer of an MLP: K = number of emulator parameters): plain FP64, one thread per output
__global__ void mlp_small_layer_kernel(const double* __restrict__ X, long long rows, int K, long long ld, const double* __restrict__ W, const double* __restrict__ bias,
                                       int ncols, int act, double* __restrict__ out, long long ldo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = idx / ncols;
  const int col = (int)(idx - row * ncols);
  if (row >= rows) return;
  double y = bias ? bias[col] : 0.;
  for (int k = 0; k < K; ++k) y = fma(X[row * ld + k], W[(size_t)k * ncols + col], y);
  if (act == 1) y = oz_silu(y);
  else if (act == 2) y = fmax(y, 0.);
  else if (act == 3) y = tanh(y);
  out[row * ldo + col] = y;
}

}  // namespace

int launch_mlp_small_layer(const double* X, long long rows, int K, long long ld, const double* W, const double* bias, int ncols, int act, double* out, long long ldo,
                           cudaStream_t stream) {
  const long long total = rows * ncols;
  mlp_small_layer_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(X, rows, K, ld, W, bias, ncols, act, out, ldo);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------------
size_t ozaki_slice_bytes(long long rows_pad, int K, int ns) { return (size_t)rows_pad * ((K + OZ_KB - 1) / OZ_KB * OZ_KB) * ns; }
long long ozaki_rows_pad(long long rows) { return (rows + OZ_M - 1) / OZ_M * OZ_M; }

// B [K, N] row-major FP64 (a layer's `kernel`) -> per-column slices in the operand image [col tiles][nkb][ns][OZ_B_SLICE] + column exponents (host, plan time)
void ozaki_slice_weights(const double* W, int K, int N, int ns, std::vector<int8_t>& slices, std::vector<int>& exps) {
  const int Kpad = (K + OZ_KB - 1) / OZ_KB * OZ_KB, nkb = Kpad / OZ_KB, ntn = (N + OZ_N - 1) / OZ_N;
  slices.assign((size_t)ntn * nkb * ns * OZ_B_SLICE, 0);
  exps.assign((size_t)ntn * OZ_N, 0);
  for (int n = 0; n < N; ++n) {
    double amax = 0.;
    for (int k = 0; k < K; ++k) amax = std::max(amax, std::fabs(W[(size_t)k * N + n]));
    int E = 0;
    if (amax > 0. && std::isfinite(amax)) { int e; std::frexp(amax, &e); E = e + 1; }
    exps[n] = E;
    const int tn = n / OZ_N, nl = n % OZ_N;
    for (int k = 0; k < K; ++k) {
      double r = std::scalbn(W[(size_t)k * N + n], -E);
      const int kb = k / OZ_KB, kl = k % OZ_KB;
      for (int s = 0; s < ns; ++s) {
        r *= 128.;
        const double qs = std::rint(r);
        r -= qs;
        slices[((size_t)(tn * nkb + kb) * ns + s) * OZ_B_SLICE + oz_tile_offset(OZ_N, nl, kl)] = (int8_t)(int)qs;
      }
    }
  }
}

// out[rows, ncols] (leading dimension ldo) = act(X[rows, K] (ld) . W + bias), W pre-sliced by ozaki_slice_weights
int launch_ozaki_layer(const double* X, long long rows, int K, long long ld, const int8_t* Bslices, const double* Sb, const double* bias, int ncols, int ns, int act,
                       double* out, long long ldo, int8_t* work_slices, int* work_exps, int nsm, cudaStream_t stream) {
  DLB_REQUIRE(ns >= 2 && ns <= OZ_MAXS, "ozaki gemm: 2 <= slices <= 8");
  const int Kpad = (K + OZ_KB - 1) / OZ_KB * OZ_KB, nkb = Kpad / OZ_KB;
  DLB_REQUIRE((long long)Kpad * (ns) * 64LL * 64LL < (1LL << 31), "ozaki gemm: contraction too long for exact int32 accumulation");
  const long long rows_pad = ozaki_rows_pad(rows);
  ozaki_slice_kernel<<<(unsigned)((rows_pad + 3) / 4), 128, 0, stream>>>(X, rows, rows_pad, K, Kpad, ld, ns, work_slices, work_exps);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  OzakiParams p;
  std::memset(&p, 0, sizeof(p));
  p.A = work_slices; p.B = Bslices; p.Ea = work_exps; p.Sb = Sb; p.bias = bias; p.out = out; p.ldo = ldo; p.rows = rows; p.ncols = ncols;
  p.nkb = nkb; p.ns = ns; p.act = act; p.ntm = (int)(rows_pad / OZ_M); p.ntn = (ncols + OZ_N - 1) / OZ_N;
  const size_t stage = (size_t)ns * (OZ_A_SLICE + OZ_B_SLICE);
  p.nstages = (int)std::min<size_t>(6, (size_t)(200 * 1024) / stage);
  DLB_REQUIRE(p.nstages >= 2, "ozaki gemm: shared memory too small for two stages");
  const size_t smem = p.nstages * stage + (size_t)(2 * p.nstages + 2) * 8 + 16;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    DLB_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long long ntiles = (long long)p.ntm * p.ntn;
  ozaki_gemm_kernel<<<(unsigned)std::min<long long>(nsm, ntiles), OZ_THREADS, smem, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb
