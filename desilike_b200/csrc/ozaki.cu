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
//   ozaki_gemm_kernel  : persistent, warp-specialised: warp 8 = TMA producer (cp.async.bulk + mbarrier ring), warp 9 = MMA issuer (one elected
//                        thread, tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 64, K = 32 per instruction, ns accumulators of 64 TMEM columns),
//                        warps 0-7 = epilogue: warp w drains TMEM lanes 32 (w & 3) .., columns 32 (w >> 2) .. of every accumulator (tcgen05.ld
//                        32x32b.x32) into two exact int64 partial sums per output, hands TMEM back to the MMA warp, and only then does the FP64
//                        work (scales, bias, activation, stores) -- which therefore overlaps the next tile's MMAs.
//   ozaki_gram_kernel  : the chi2 / normal-equation reduction over the columns of a GEMM output D[point, term, column] (the contraction of a
//                        likelihood block with wide MLP features run as ONE GEMM over all its terms): slot values from the per-point bias
//                        coefficients, pairwise dot products, Gram partials in the layout of contract_kernel's epilogue.
#include "common.h"

#include <cstring>
#include <cstdio>

namespace dlb {

namespace {

constexpr int OZ_M = 128, OZ_N = 64, OZ_KB = 32;          // output tile, k-block (bytes = int8 elements) per pipeline stage = one MMA K
constexpr int OZ_A_SLICE = OZ_M * OZ_KB, OZ_B_SLICE = OZ_N * OZ_KB;   // bytes of one slice of one stage
constexpr int OZ_MAXS = 8;
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_MMA_WARPS = 2;
constexpr int OZ_THREADS = (OZ_EPI_WARPS + 1 + OZ_MMA_WARPS + 1) * 32;   // warps 0-7 epilogue, 8 TMA producer, 9-10 MMA issuers, 11 idle (12 warps: 3 per scheduler)

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

// ---- slicing: one block of 8 warps per 8-row group (warp w -> row w of the group).  Rows are read coalesced; the int8 slices are staged in shared
//      memory in the operand image's order, so that the block writes 128 contiguous bytes (8 rows x 16 bytes of one k chunk) per 8 lanes ----
__global__ void __launch_bounds__(256) ozaki_slice_kernel(const double* __restrict__ X, long long rows, long long rows_pad, int K, int Kpad, long long ld, int ns,
                                                          int8_t* __restrict__ slices, int* __restrict__ exps) {
  extern __shared__ __align__(16) unsigned char sl_smem[];   // [ns][Kpad / 16][8 rows][16 bytes]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;    // rows_pad is a multiple of OZ_M
  const int nkb = Kpad / OZ_KB, nc16 = Kpad >> 4;
  const double* x = X + row * ld;
  double amax = 0.;
  if (row < rows)
    for (int k = lane; k < K; k += 32) amax = fmax(amax, fabs(x[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  int E = 0;
  if (amax > 0. && isfinite(amax)) { int e; frexp(amax, &e); E = e + 1; }   // amax = m 2^e, m in [0.5, 1)  ->  amax 2^-(e + 1) < 1/2
  if (lane == 0) exps[row] = E;
  const double sc = scalbn(1., -E);
  const size_t sstride = (size_t)Kpad * 8;
  // four consecutive k per lane: 32-byte loads (rows 32-byte aligned: `vec`), one 32-bit shared-memory store per slice
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<unsigned long long>(X) & 31ull) == 0ull;
  for (int k = 4 * lane; k < Kpad; k += 128) {
    double r[4];
    if (row < rows && vec && k + 4 <= K) {
      const double4 v = *reinterpret_cast<const double4*>(x + k);
      r[0] = v.x * sc; r[1] = v.y * sc; r[2] = v.z * sc; r[3] = v.w * sc;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (row < rows && k + j < K) ? x[k + j] * sc : 0.;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) if (!isfinite(r[j])) r[j] = 0.;
    unsigned char* dst = sl_smem + (size_t)(k >> 4) * 128 + warp * 16 + (k & 15);
#pragma unroll
    for (int s = 0; s < OZ_MAXS; ++s)
      if (s < ns) {
        uint32_t w = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r[j] *= 128.;
          const double qs = rint(r[j]);
          r[j] -= qs;
          w |= ((uint32_t)(int)qs & 0xFFu) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)s * sstride) = w;
      }
  }
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * 8, rt = r0 / OZ_M;
  const int g8 = (int)(r0 - rt * OZ_M) >> 3;
  for (int u = threadIdx.x; u < ns * nc16 * 8; u += 256) {
    const int r8 = u & 7, sc_ = u >> 3, s = sc_ / nc16, c = sc_ - s * nc16;
    const int kb = c / (OZ_KB / 16), kc = c % (OZ_KB / 16);
    int8_t* dst = slices + ((size_t)(rt * nkb + kb) * ns + s) * OZ_A_SLICE + (size_t)kc * (OZ_M * 16) + (size_t)g8 * 128 + (size_t)r8 * 16;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(sl_smem + (size_t)u * 16);
  }
}

struct OzakiParams {
  const int8_t* A;        // [row tiles][nkb][ns][OZ_A_SLICE]
  const int8_t* B;        // [col tiles][nkb][k chunk of 16][ns][64 columns x 16 bytes]   (plan constant; the slices side by side along N)
  const int* Ea;          // [rows_pad]
  const double* Sb;       // [ncols_pad] 2^Eb as doubles (epilogue)
  const double* bias;     // [ncols] or NULL
  double* out; long long ldo; long long rows; int ncols;
  int nkb, ns, act, ntm, ntn;
  int nstages;
  int C, ntg;             // thread-block cluster size along N (1, 2 or 4) and column-tile groups per row tile = ceil(ntn / C)
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE; LBO = byte distance between the two 16-byte k chunks of one
// MMA, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ double oz_silu(double v) { return v * fast_rcp(1. + fast_exp(fmin(fmax(-v, -708.), 708.))); }

// All slice products of one k block: acc_t += A_i . B_j^T for i + j = t <= NS - 1.  The B slices of a stage lie side by side along N ([k chunk][slice]
// [64 rows x 16 bytes]) and accumulator t sits at TMEM column 64 t, so ONE tcgen05.mma of N = 64 m multiplies A_i with the m consecutive slices
// B_j .. B_j+m-1 into the accumulators i + j .. i + j + m - 1 (m <= 4: N <= 256).
// Issuing one tcgen05.mma costs a single warp ~100 cycles (ELECT -> R2UR x 4 -> UTCIMMA, a dependent chain) while a 128 x 64 x 32 product occupies the
// tensor pipe for 32: TWO warps issue, each owning a disjoint set of accumulators (PART 0: t < T0, PART 1: t >= T0, T0 = (2 NS + 1) / 3 balances their
// instruction counts) -- no two warps ever write the same TMEM columns, and every warp's own instructions stay in program order.
// `accumulate` is false for the first k block of a tile, where the i = 0 products overwrite every accumulator.  Executed by the whole (converged)
// warp with warp-uniform operands; `elected` predicates the instruction itself.
template <int NS, int PART>
__device__ __forceinline__ void oz_issue_kblock(uint32_t a0, uint32_t b0, uint32_t tmem, bool accumulate, uint32_t elected) {
  constexpr int T0 = (2 * NS + 1) / 3;
  const uint64_t da0 = oz_desc(a0, OZ_M * 16, 128), db0 = oz_desc(b0, NS * (OZ_N * 16), 128);   // (the 14-bit address fields cannot carry: shared memory < 256 KB)
  const uint32_t acc0 = accumulate ? 1u : 0u;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const uint64_t da = da0 + (uint64_t)((i * OZ_A_SLICE) >> 4);
    // B slices j with i + j in this warp's accumulator range
    const int jlo = PART == 0 ? 0 : (T0 - i > 0 ? T0 - i : 0), jhi = PART == 0 ? (T0 - i < NS - i ? T0 - i : NS - i) : NS - i;   // [jlo, jhi)
#pragma unroll
    for (int j0 = jlo; j0 < jhi; j0 += 4) {
      const int m = jhi - j0 < 4 ? jhi - j0 : 4;
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((OZ_N * m) >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
      const uint64_t db = db0 + (uint64_t)((j0 * (OZ_N * 16)) >> 4);
      const uint32_t d = tmem + (uint32_t)((i + j0) * OZ_N);
      asm volatile(
          "{\n\t.reg .pred p, q;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "setp.ne.b32 q, %6, 0;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
          ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i == 0 ? acc0 : 1u), "r"(0u), "r"(elected) : "memory");
    }
  }
}

template <int PART>
__device__ __forceinline__ void oz_issue_kblock_ns(int ns, uint32_t a0, uint32_t b0, uint32_t tmem, bool accumulate, uint32_t elected) {
  switch (ns) {   // fully unrolled per slice count
    case 2: oz_issue_kblock<2, PART>(a0, b0, tmem, accumulate, elected); break;
    case 3: oz_issue_kblock<3, PART>(a0, b0, tmem, accumulate, elected); break;
    case 4: oz_issue_kblock<4, PART>(a0, b0, tmem, accumulate, elected); break;
    case 5: oz_issue_kblock<5, PART>(a0, b0, tmem, accumulate, elected); break;
    case 6: oz_issue_kblock<6, PART>(a0, b0, tmem, accumulate, elected); break;
    case 7: oz_issue_kblock<7, PART>(a0, b0, tmem, accumulate, elected); break;
    default: oz_issue_kblock<8, PART>(a0, b0, tmem, accumulate, elected); break;
  }
}

// 32 consecutive TMEM columns of this thread's lane -> registers
__device__ __forceinline__ void oz_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int ACT>
__device__ __forceinline__ double oz_act(double y) {
  if constexpr (ACT == 1) return oz_silu(y);
  else if constexpr (ACT == 2) return fmax(y, 0.);
  else if constexpr (ACT == 3) return tanh(y);
  else return y;
}

// ACT: activation of the epilogue (0 none, 1 silu, 2 relu, 3 tanh; reference emulators/conversion.py:27-31)
template <int ACT>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ OzakiParams p) {
  extern __shared__ __align__(1024) unsigned char oz_smem[];
  const int ns = p.ns, S = p.nstages, C = p.C;
  const uint32_t stage_bytes = (uint32_t)ns * (OZ_A_SLICE + OZ_B_SLICE);
  unsigned char* stages = oz_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + (size_t)S * stage_bytes);   // full[S], empty[S], tmem_full, tmem_empty
  // Thread-block cluster of C CTAs along N: the C CTAs work on the column tiles tg C .. tg C + C - 1 of the SAME row tile in lock step and share
  // its A slices -- CTA r fetches the r-th part of every A stage and multicasts it into the shared memory of all C CTAs (cp.async.bulk
  // .multicast::cluster).  Opt-in (C = 1 by default, see launch_gemm_act for the measurements).  A stage is released to the producers when the MMAs of ALL C CTAs have read it: every MMA
  // warp commits to the stage's `empty` barrier of every CTA (tcgen05.commit .multicast::cluster).
  uint32_t crank = 0;
  if (C > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = oz_smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t tmem_full = bar0 + 8u * (2 * S), tmem_empty = bar0 + 8u * (2 * S + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full(s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty(s)), "r"(C * OZ_MMA_WARPS));   // every MMA warp of every CTA of the cluster reads (the multicast A chunks of) a stage
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tmem_full), "r"(OZ_MMA_WARPS));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tmem_empty), "r"(OZ_EPI_WARPS * 32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (one CTA per SM); the allocating warp also frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(oz_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (C > 1) {   // no CTA may multicast into / arrive on a peer whose barriers are not initialised yet
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // work units: (row tile, group of C column tiles); cluster c takes units c, c + nclusters, ...; CTA r of the cluster the column tile tg C + r
  // (beyond ntn: a dummy tile -- the CTA still takes part in the multicast loads and barriers, its results are not stored)
  const long long nunits = (long long)p.ntm * p.ntg, unit0 = blockIdx.x / C, ustep = gridDim.x / C;
  if (warp == OZ_EPI_WARPS) {
    // ===== producer: one bulk copy of all A slices + one of all B slices per (tile, k block) =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const uint32_t a_bytes = (uint32_t)ns * OZ_A_SLICE, a_part = a_bytes / (uint32_t)C;
      for (long long unit = unit0; unit < nunits; unit += ustep) {
        const long long tm = unit / p.ntg;
        const int tn = min((int)(unit - tm * p.ntg) * C + (int)crank, p.ntn - 1);
        for (int kb = 0; kb < p.nkb; ++kb) {
          oz_mbar_wait(empty(s), ph ^ 1u);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full(s)), "r"(stage_bytes) : "memory");
          const int8_t* srcA = p.A + ((size_t)(tm * p.nkb + kb) * ns) * OZ_A_SLICE;
          const int8_t* srcB = p.B + ((size_t)((long long)tn * p.nkb + kb) * ns) * OZ_B_SLICE;
          const uint32_t dst = oz_smem_u32(stages + (size_t)s * stage_bytes);
          if (C > 1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(dst + crank * a_part), "l"(srcA + (size_t)crank * a_part), "r"(a_part), "r"(full(s)), "h"(cmask) : "memory");
          else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(srcA), "r"(a_bytes), "r"(full(s))
                         : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + a_bytes), "l"(srcB),
                       "r"((uint32_t)ns * OZ_B_SLICE), "r"(full(s)) : "memory");
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp > OZ_EPI_WARPS && warp <= OZ_EPI_WARPS + OZ_MMA_WARPS) {
    // ===== MMA issuers (two warps, disjoint accumulators): the whole warp walks the pipeline (warp-uniform control flow and operands), lane 0 issues =====
    const int part = warp - (OZ_EPI_WARPS + 1);
    const uint32_t elected = lane == 0 ? 1u : 0u;
    int s = 0; uint32_t ph = 0, tph = 0;
    for (long long unit = unit0; unit < nunits; unit += ustep) {
      oz_mbar_wait(tmem_empty, tph ^ 1u);   // the epilogue has drained the accumulators of the previous tile
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < p.nkb; ++kb) {
        oz_mbar_wait(full(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = oz_smem_u32(stages) + (uint32_t)s * stage_bytes, b0 = a0 + (uint32_t)ns * OZ_A_SLICE;
        if (part == 0) oz_issue_kblock_ns<0>(ns, a0, b0, tmem, kb > 0, elected);
        else oz_issue_kblock_ns<1>(ns, a0, b0, tmem, kb > 0, elected);
        // the stage may be refilled once these MMAs have read it
        if (elected) {
          if (C > 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(empty(s)), "h"(cmask) : "memory");
          else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty(s)) : "memory");
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tmem_full) : "memory");
      __syncwarp();
      tph ^= 1u;
    }
  } else if (warp < OZ_EPI_WARPS) {
    // ===== epilogue: warp w reads TMEM lanes 32 (w & 3) .. + 31 (= output rows of the tile), columns 32 (w >> 2) .. + 31 of every accumulator =====
    const int q = warp & 3, h = warp >> 2;
    const int gs = (ns + 1) >> 1;                     // accumulators per exact int64 group: (t + 1) K 64^2 2^(7 (gs - 1)) < 2^46 for gs <= 4
    const double w0 = scalbn(1., -7 * (gs + 1)), w1 = scalbn(1., -7 * (2 * gs + 1));
    uint32_t tph = 0;
    for (long long unit = unit0; unit < nunits; unit += ustep) {
      const long long tm = unit / p.ntg;
      const int tn = (int)(unit - tm * p.ntg) * C + (int)crank;   // >= ntn: dummy tile
      oz_mbar_wait(tmem_full, tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long row = tm * OZ_M + q * 32 + lane;
      // ---- phase A (TMEM busy): the int32 accumulators t = 0 .. gs - 1 and t = gs .. ns - 1 are combined EXACTLY in int64,
      //      G_0 = sum_u acc_u 2^(7 (gs - 1 - u)),  G_1 = sum_u acc_(gs + u) 2^(7 (gs - 1 - u));  v = G_1 2^(-7 (2 gs + 1)) + G_0 2^(-7 (gs + 1)) ----
      long long G0[32], G1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) G0[j] = G1[j] = 0;
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32);
      for (int t = 0; t < gs; ++t) {
        uint32_t r[32];
        oz_tmem_ld32(tbase + (uint32_t)(t * OZ_N), r);
        const int sh = 7 * (gs - 1 - t);
#pragma unroll
        for (int j = 0; j < 32; ++j) G0[j] += (long long)((unsigned long long)(long long)(int)r[j] << sh);
      }
      for (int t = gs; t < ns; ++t) {
        uint32_t r[32];
        oz_tmem_ld32(tbase + (uint32_t)(t * OZ_N), r);
        const int sh = 7 * (2 * gs - 1 - t);
#pragma unroll
        for (int j = 0; j < 32; ++j) G1[j] += (long long)((unsigned long long)(long long)(int)r[j] << sh);
      }
      // TMEM is free again: the MMA warp starts the next tile while this warp finishes its rows
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty) : "memory");
      // ---- phase B: exponents, bias, activation, store -- 8 outputs at a time: their column scales / biases come through the read-only path
      //      (no aliasing with the stores, so the loads are issued up front) and the 8 FP64 chains interleave ----
      if (row < p.rows && tn < p.ntn) {
        const double sa = scalbn(1., p.Ea[row]);
        const int col0 = tn * OZ_N + h * 32;
        double* orow = p.out + row * p.ldo + col0;
        const double* Sb = p.Sb + col0;
        const double* bias = p.bias ? p.bias + col0 : nullptr;
        const bool full = col0 + 32 <= p.ncols, al16 = (reinterpret_cast<unsigned long long>(orow) & 15ull) == 0ull;
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
          double sb[8], bi[8], y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { sb[j] = __ldg(Sb + j0 + j); bi[j] = bias ? __ldg(bias + j0 + j) : 0.; }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double v = fma((double)G0[j0 + j], w0, (double)G1[j0 + j] * w1);
            y[j] = oz_act<ACT>(v * sa * sb[j] + bi[j]);
          }
          if (full && al16) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(orow + j0 + j) = make_double2(y[j], y[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (full || col0 + j0 + j < p.ncols) orow[j0 + j] = y[j];
          }
        }
      }
      tph ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (C > 1) {   // peers may still multicast into this CTA's shared memory / arrive on its barriers until they are done, too
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
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

// ---- chi2 / normal-equation reduction of a GEMM output -------------------------------------------------------------------------------------
// D[point, t * ncolx + col] = (features . M_t)[col] for the nterms terms of one likelihood block; columns ncols .. ncolx - 1 are the auxiliary
// columns M_t g_i (dot products with the constant whitened Jacobian rows, csrc/plan.cu: build_block_core).  Slot values
//     S_a[col] = sum_{t in slot a} coef_t(theta) D_t[col] - (a == 0) r0[col]
// then the Gram partials in the order of contract_kernel's epilogue: S_a . S_b (a <= b) over the data columns, followed by S_a[ncols + i].
// One warp per point (lanes stride over the columns: coalesced reads), OZ_GRAM_WARPS points per block; results staged in shared memory so
// that the point-minor partial rows are written in contiguous runs.
constexpr int OZ_GRAM_WARPS = 8, OZ_GRAM_RPW = 4, OZ_GRAM_ROWS = OZ_GRAM_WARPS * OZ_GRAM_RPW;   // rows per warp / per block

template <int R>
__global__ void __launch_bounds__(OZ_GRAM_WARPS * 32, R <= 4 ? 4 : 2) ozaki_gram_kernel(const OzakiGramParams p) {
  constexpr int NP = R * (R + 1) / 2;
  __shared__ double red[OZ_GRAM_ROWS][NP + R * 8 + 1];
  __shared__ double cs[OZ_GRAM_WARPS][OZ_GRAM_MAXT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * OZ_GRAM_ROWS;
  const int Rb = p.R, nS = Rb * (Rb + 1) / 2, npart = nS + Rb * p.nconst;
  for (int rr = 0; rr < OZ_GRAM_RPW; ++rr) {
    const int rl = rr * OZ_GRAM_WARPS + warp;
    const long long row = row0 + rl;
    if (row >= p.N) break;
    __syncwarp();
    if (lane < p.nterms) {   // bias monomials of this point, one lane per term
      double v = 1.;
      if (p.tcoef[lane] >= 0) {
        const DevCoef* cf = p.coefs + p.tcoef[lane];
        v = cf->scale;
        for (int f = 0; f < cf->nf; ++f) {
          const double x = p.params[row * p.d + cf->idx[f]];
          const int e = cf->pow[f];
          double pw = 1.;
          for (int i = 0; i < (e >= 0 ? e : -e); ++i) pw *= x;
          v *= e >= 0 ? pw : 1. / pw;
        }
      }
      cs[warp][lane] = v;
    }
    __syncwarp();
    const double* Drow = p.D + row * p.ldD;
    double acc[NP];
#pragma unroll
    for (int e = 0; e < NP; ++e) acc[e] = 0.;
    // two columns per lane and step (ncolx and ldD are even: 16-byte loads); column ncols (odd ncols) belongs to the auxiliary columns: masked
    for (int col = 2 * lane; col < p.ncols; col += 64) {
      double S0[R], S1[R];
#pragma unroll
      for (int a = 0; a < R; ++a) S0[a] = S1[a] = 0.;
      const double2 r0v = *reinterpret_cast<const double2*>(p.r0 + col);
      S0[0] = -r0v.x; S1[0] = -r0v.y;
#pragma unroll 6
      for (int t = 0; t < p.nterms; ++t) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(Drow + (size_t)t * p.ncolx + col));   // streamed once
        const double c = cs[warp][t];
        const int sl = p.tslot[t];
#pragma unroll
        for (int a = 0; a < R; ++a) if (sl == a) { S0[a] = fma(c, v.x, S0[a]); S1[a] = fma(c, v.y, S1[a]); }
      }
      if (col + 1 >= p.ncols) {
#pragma unroll
        for (int a = 0; a < R; ++a) S1[a] = 0.;
      }
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int b = a; b < R; ++b) {
          const int e = a * R - (a * (a - 1)) / 2 + (b - a);
          acc[e] = fma(S0[a], S0[b], acc[e]);
          acc[e] = fma(S1[a], S1[b], acc[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < NP; ++e) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int b = a; b < R; ++b)
          if (b < Rb) red[rl][a * Rb - (a * (a - 1)) / 2 + (b - a)] = acc[a * R - (a * (a - 1)) / 2 + (b - a)];
    }
    // auxiliary columns: lane = (i, a)
    for (int e = lane; e < Rb * p.nconst; e += 32) {
      const int i = e / Rb, a = e - i * Rb, col = p.ncols + i;
      double v = a == 0 ? -p.r0[col] : 0.;
      for (int t = 0; t < p.nterms; ++t)
        if (p.tslot[t] == a) v += cs[warp][t] * Drow[(size_t)t * p.ncolx + col];
      red[rl][nS + e] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npart * OZ_GRAM_ROWS; i += OZ_GRAM_WARPS * 32) {
    const int e = i / OZ_GRAM_ROWS, w = i - e * OZ_GRAM_ROWS;
    if (row0 + w < p.N) p.partial[(p.poff + e) * p.npad + row0 + w] = red[w][e];
  }
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

// B [K, N] row-major FP64 (a layer's `kernel`) -> per-column slices in the operand image [col tiles][nkb][k chunk][ns][64 x 16 bytes] + column exponents (host, plan time)
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
        slices[(size_t)(tn * nkb + kb) * ns * OZ_B_SLICE + (size_t)(kl >> 4) * (ns * OZ_N * 16) + (size_t)s * (OZ_N * 16) + (size_t)(nl >> 3) * 128 + (size_t)(nl & 7) * 16 + (kl & 15)] = (int8_t)(int)qs;
      }
    }
  }
}

template <int ACT>
static int launch_gemm_act(OzakiParams& p, int nsm, size_t smem, cudaStream_t stream) {
  static bool configured_dev[64] = {};
  static int max_clusters_dev[64][9] = {};
  const int slot = current_device_slot();
  if (!configured_dev[slot]) {
    DLB_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_dev[slot] = true;
  }
  // Cluster size.  Measured on B200 (c4wide, 16 384 points, 6 slices): 1 -> 2.86e6, 2 -> 2.83e6, 4 -> 2.64e6, 8 -> 2.46e6 evals/s (only 15 clusters of 8 =
  // 120 SMs are co-schedulable): L2 already merges the concurrent reads of one A row tile by neighbouring CTAs, so the multicast saves no L2 -> SM
  // bandwidth while the lock step of the cluster costs.  Default 1; DLB_OZAKI_CLUSTER = 2 | 4 | 8 keeps the multicast path testable.
  int C = 1;
  if (const char* e = getenv("DLB_OZAKI_CLUSTER")) { const int c = atoi(e); if ((c == 1 || c == 2 || c == 4 || c == 8) && p.ntn >= c) C = c; }
  p.C = C; p.ntg = (p.ntn + C - 1) / C;
  const long long nunits = (long long)p.ntm * p.ntg;
  if (C > 1 && max_clusters_dev[slot][C] == 0) {   // cluster size forced through the environment: query its occupancy, too
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3((unsigned)(nsm / C * C)); q.blockDim = dim3(OZ_THREADS); q.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = (unsigned)C; at.val.clusterDim.y = at.val.clusterDim.z = 1;
    q.attrs = &at; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ozaki_gemm_kernel<ACT>, &q) != cudaSuccess) { cudaGetLastError(); n = -1; }
    max_clusters_dev[slot][C] = n > 0 ? n : -1;
    if (getenv("DLB_OZAKI_VERBOSE")) fprintf(stderr, "[ozaki] cluster size %d: %d active clusters\n", C, n);
  }
  int nclusters = C > 1 ? max_clusters_dev[slot][C] : nsm;
  if (nclusters <= 0) nclusters = nsm / C;
  const unsigned grid = (unsigned)(std::min<long long>(nclusters, nunits) * C);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = (unsigned)C; at.val.clusterDim.y = at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  DLB_CUDA(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<ACT>, p));
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

// rows of X[rows, K] (leading dimension ld) -> int8 slices (operand image) + row exponents
int launch_ozaki_slice(const double* X, long long rows, int K, long long ld, int ns, int8_t* work_slices, int* work_exps, cudaStream_t stream) {
  DLB_REQUIRE(ns >= 2 && ns <= OZ_MAXS, "ozaki gemm: 2 <= slices <= 8");
  const int Kpad = (K + OZ_KB - 1) / OZ_KB * OZ_KB;
  const long long rows_pad = ozaki_rows_pad(rows);
  const size_t smem = (size_t)ns * Kpad * 8;
  DLB_REQUIRE(smem <= (size_t)160 * 1024, "ozaki slicing: contraction too long for the shared-memory staging");
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    DLB_CUDA(cudaFuncSetAttribute(ozaki_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  ozaki_slice_kernel<<<(unsigned)(rows_pad / 8), 256, smem, stream>>>(X, rows, rows_pad, K, Kpad, ld, ns, work_slices, work_exps);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

// out[rows, ncols] (leading dimension ldo) = act((sliced A) . W + bias), A pre-sliced by launch_ozaki_slice, W by ozaki_slice_weights
int launch_ozaki_gemm(long long rows, int K, const int8_t* Aslices, const int* Ea, const int8_t* Bslices, const double* Sb, const double* bias, int ncols, int ns, int act,
                      double* out, long long ldo, int nsm, cudaStream_t stream) {
  DLB_REQUIRE(ns >= 2 && ns <= OZ_MAXS, "ozaki gemm: 2 <= slices <= 8");
  const int Kpad = (K + OZ_KB - 1) / OZ_KB * OZ_KB, nkb = Kpad / OZ_KB;
  DLB_REQUIRE((long long)Kpad * (ns) * 64LL * 64LL < (1LL << 31), "ozaki gemm: contraction too long for exact int32 accumulation");
  const long long rows_pad = ozaki_rows_pad(rows);
  OzakiParams p;
  std::memset(&p, 0, sizeof(p));
  p.A = Aslices; p.B = Bslices; p.Ea = Ea; p.Sb = Sb; p.bias = bias; p.out = out; p.ldo = ldo; p.rows = rows; p.ncols = ncols;
  p.nkb = nkb; p.ns = ns; p.act = act; p.ntm = (int)(rows_pad / OZ_M); p.ntn = (ncols + OZ_N - 1) / OZ_N;
  const size_t stage = (size_t)ns * (OZ_A_SLICE + OZ_B_SLICE);
  p.nstages = (int)std::min<size_t>(8, (size_t)(200 * 1024) / stage);
  DLB_REQUIRE(p.nstages >= 2, "ozaki gemm: shared memory too small for two stages");
  const size_t smem = p.nstages * stage + (size_t)(2 * p.nstages + 2) * 8 + 16;
  switch (act) {
    case 0: return launch_gemm_act<0>(p, nsm, smem, stream);
    case 1: return launch_gemm_act<1>(p, nsm, smem, stream);
    case 2: return launch_gemm_act<2>(p, nsm, smem, stream);
    case 3: return launch_gemm_act<3>(p, nsm, smem, stream);
  }
  set_error("ozaki gemm: unknown activation");
  return DLB_ERR_ARG;
}

// out[rows, ncols] (leading dimension ldo) = act(X[rows, K] (ld) . W + bias), W pre-sliced by ozaki_slice_weights
int launch_ozaki_layer(const double* X, long long rows, int K, long long ld, const int8_t* Bslices, const double* Sb, const double* bias, int ncols, int ns, int act,
                       double* out, long long ldo, int8_t* work_slices, int* work_exps, int nsm, cudaStream_t stream) {
  DLB_TRY(launch_ozaki_slice(X, rows, K, ld, ns, work_slices, work_exps, stream));
  return launch_ozaki_gemm(rows, K, work_slices, work_exps, Bslices, Sb, bias, ncols, ns, act, out, ldo, nsm, stream);
}

int launch_ozaki_gram(const OzakiGramParams& p, cudaStream_t stream) {
  DLB_REQUIRE(p.nterms >= 1 && p.nterms <= OZ_GRAM_MAXT && p.R >= 1 && p.R <= 8 && p.nconst >= 0 && p.nconst <= 8, "ozaki gram: 1 <= terms <= 16, 1 <= slots <= 8, nconst <= 8");
  DLB_REQUIRE((p.ncolx & 1) == 0 && (p.ldD & 1) == 0, "ozaki gram: even column counts required (16-byte loads)");
  const unsigned grid = (unsigned)((p.N + OZ_GRAM_ROWS - 1) / OZ_GRAM_ROWS);
  if (p.R <= 1) ozaki_gram_kernel<1><<<grid, OZ_GRAM_WARPS * 32, 0, stream>>>(p);
  else if (p.R <= 2) ozaki_gram_kernel<2><<<grid, OZ_GRAM_WARPS * 32, 0, stream>>>(p);
  else if (p.R <= 4) ozaki_gram_kernel<4><<<grid, OZ_GRAM_WARPS * 32, 0, stream>>>(p);
  else ozaki_gram_kernel<8><<<grid, OZ_GRAM_WARPS * 32, 0, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb
