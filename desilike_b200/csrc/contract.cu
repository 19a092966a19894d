// Persistent fused contraction kernel (sm_100a).  For every tile of parameter points and every output block (a group of
// data-vector columns, one per tracer when the precision is block-diagonal)
//
//     slot[s][p, :] = sum_{terms t of slot s} coef_t[p] * A[p, fcol_t : fcol_t + K_t] . M_t[K_t, cols]
//
// followed, in registers, by the chi2 / normal-equation (Gram) reduction over the columns.  This one kernel replaces,
// per point, the reference's   emulator predict (Taylor monomials + tensordot) -> tracer bias combine -> window matrix apply
// -> flatdiff -> d^T P d   and the J P d, J P J^T products of the analytic marginalisation
//   (reference emulators/__init__.py:471-507; full_shape.py:545-550,628-634; window.py:459-473; likelihoods/base.py:13-17,
//    :165-174, :658-664): all constant matrices (Taylor derivatives / last MLP layer, window W, Cholesky factor L of the
//    precision) are folded into M_t at plan time, so no theory vector ever touches HBM.
//
// Hardware mapping
//   * one persistent CTA per SM; the batch is split into row units (WM * 8 rows) and every CTA owns a contiguous range of
//     units (balanced to one unit, so there is no wave-quantisation tail); it walks its range in tiles of up to MT units
//     and, per tile, over all blocks / column blocks / k-chunks.  One launch covers the whole likelihood.
//   * FP64 tensor pipe: DMMA.8x8x4 (mma.sync.m8n8k4.f64) -- tcgen05 has no FP64 kind; every FP64 mma shape lowers to
//     DMMA.8x8x4 on sm_100a (checked with cuobjdump).  One warp can issue a DMMA every 32 cycles while the pipe retires one
//     per 16 cycles and scheduler (measured: dlb_measure_fp64_peak(100 + warps)), so 2 warps per scheduler have zero slack:
//     the kernel runs 3-4 warps per scheduler (12-16 warps, all of them consumers) with a (MT*8) x (NT*8) tile each,
//     warp grid WM (m) x 2 (n).
//   * B operand (pre-tiled term matrices, L2 resident) streamed with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) into an
//     mbarrier full/empty ring; there is no dedicated producer warp: warp 0 re-arms the stage that was released one chunk
//     earlier after finishing its own chunk (S - 1 chunks of prefetch).  Rows are KC = 28 or 36 doubles (= 4 mod 8 makes the
//     64-bit fragment loads bank-conflict free).  The per-column-block epilogue constants (r0, constant Jacobian rows)
//     travel through the same ring as one extra chunk.
//   * A operand, RESIDENT mode (Taylor emulators): the warps evaluate the monomials prod_j (theta_j - c_j)^p_tj of their
//     tile straight into shared memory (one warp per row, one lane per monomial) -- the feature matrix never exists in
//     HBM -- and it stays resident for all terms and column blocks.  STREAMING mode (MLP / Kaiser / BAO features computed
//     by their own kernels): A tiles are TMA-copied per chunk.
//   * the parameter tile of the NEXT row tile is prefetched with cp.async while the current one computes; bias-monomial
//     coefficients are evaluated per tile into shared memory and applied to the A fragment in registers.
//   * the Gram epilogue is warp-shuffle reductions accumulated in a per-warp-column shared-memory copy (no atomics), summed
//     over column blocks in the kernel and written once per (tile, block).
// This file is compiled twice: as is (kernel instances without the masked sub-tile path) and through contract_masked.cu
// (DLB_CONTRACT_MASKED = 1: instances with it) -- two translation units, so the two sets of variants compile in parallel.
#include "common.h"

#ifndef DLB_CONTRACT_MASKED
#define DLB_CONTRACT_MASKED 0
#endif

namespace dlb {

namespace {

constexpr bool kMasked = DLB_CONTRACT_MASKED != 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {   // bar: shared-window address computed once per kernel
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// v / (1 + exp(-v))  (reference emulators/conversion.py:27)
__device__ __forceinline__ double silu(double v) { return v * fast_rcp(1. + fast_exp(fmin(fmax(-v, -708.), 708.))); }

__device__ __forceinline__ double ipow(double x, int n) {
  double r = 1.;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// One k-chunk of DMMAs for MTC active m-tiles: acc[mt][nt] += (beta[mt] * A[mt-rows, k]) . B[k, nt-cols].
// The two warps of a pair interleave their n-tiles (warp wn owns the 8-column tiles 2 nt + wn of the column block), so that column
// blocks whose right part is structurally zero (triangular rank-reduced problems, padding of the last block) stay balanced over the pair.
// BETA = false: the A fragment is used as is (quadratic-form blocks: no per-point coefficient on the terms -- and no scalar FP64 multiply queued behind
// the DMMAs in the shared FP64 pipe).
template <int MT, int NT, int MTC, int KC, bool BETA = true>
__device__ __forceinline__ void mma_chunk(double (&acc)[MT][NT][2], const double* __restrict__ A, int astride, const double* __restrict__ B,
                                          const double (&beta)[MT], int nk4) {
#pragma unroll(BETA ? 3 : 8)   // (quadratic-form chunks are 8 k4 steps for 64 features: no remainder loop)
  for (int k4 = 0; k4 < nk4; ++k4) {
    double a[MTC], b[NT];
#pragma unroll
    for (int mt = 0; mt < MTC; ++mt) a[mt] = BETA ? A[mt * astride + k4 * 4] * beta[mt] : A[mt * astride + k4 * 4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = B[nt * 16 * KC + k4 * 4];
#pragma unroll
    for (int mt = 0; mt < MTC; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

// A run of k4 steps restricted to the first NTC n-tiles of the warp.  FULLM: every m-tile of the warp is active (the usual case: only the last
// row tile of a CTA's range is short) -- no per-DMMA predicates (predicated-off DMMAs are not free).
template <int MT, int NT, int NTC, int KC, bool FULLM>
__device__ __forceinline__ void mma_run(double (&acc)[MT][NT][2], const double* __restrict__ A, int astride, const double* __restrict__ B,
                                        const double (&beta)[MT], int k4a, int k4b, int mtc) {
#pragma unroll 2
  for (int k4 = k4a; k4 < k4b; ++k4) {
    double a[MT], b[NTC];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) a[mt] = (FULLM || mt < mtc) ? A[mt * astride + k4 * 4] * beta[mt] : 0.;
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) b[nt] = B[nt * 16 * KC + k4 * 4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (FULLM || mt < mtc) {
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
  }
}

// Masked k-chunk: `mask` lists up to 8 runs of k4 steps, 8 bits each from the low end: bits 0-4 = run length (0 terminates), bits 5-7 = number of
// leading n-tiles of this warp whose B sub-tiles (4 rows x 8 columns) are not structurally zero during the run (plan-time table,
// csrc/plan.cu: build_block_core).  Zero sub-tiles of lower-triangular (rank-reduced) term matrices and of the padding of the last column
// block are neither loaded nor multiplied.  (Measured and rejected: one code path with the loads / DMMAs predicated on nt < nact instead of
// one instantiation per count -- C5 0.90 vs 0.75 ms: predicated-off DMMAs are not free.)
template <int MT, int NT, int KC, bool FULLM>
__device__ __forceinline__ void mma_chunk_masked_m(double (&acc)[MT][NT][2], const double* A, int astride, const double* B, const double (&beta)[MT],
                                                   int mtc, unsigned long long mask) {
  int k4 = 0;
  for (;;) {
    const int len = (int)(mask & 31ull);
    if (len == 0) break;
    const int nact = (int)((mask >> 5) & 7ull), k4e = k4 + len;
    mask >>= 8;
    if (nact >= NT) mma_run<MT, NT, NT, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc);
    else switch (nact) {
      case 0: break;
      case 1: mma_run<MT, NT, 1, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
      case 2: if constexpr (NT > 2) mma_run<MT, NT, 2, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
      case 3: if constexpr (NT > 3) mma_run<MT, NT, 3, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
      case 4: if constexpr (NT > 4) mma_run<MT, NT, 4, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
      case 5: if constexpr (NT > 5) mma_run<MT, NT, 5, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
      default: if constexpr (NT > 6) mma_run<MT, NT, 6, KC, FULLM>(acc, A, astride, B, beta, k4, k4e, mtc); break;
    }
    k4 = k4e;
  }
}

template <int MT, int NT, int KC>
__device__ __forceinline__ void mma_chunk_masked(double (&acc)[MT][NT][2], const double* A, int astride, const double* B, const double (&beta)[MT],
                                                 int mtc, unsigned long long mask) {
  if (mtc == MT) mma_chunk_masked_m<MT, NT, KC, true>(acc, A, astride, B, beta, mtc, mask);
  else mma_chunk_masked_m<MT, NT, KC, false>(acc, A, astride, B, beta, mtc, mask);
}

template <int MT, int NT, int KC, bool MASKED, bool BETA = true>
__device__ __forceinline__ void mma_chunk_dispatch(double (&acc)[MT][NT][2], const double* A, int astride, const double* B, const double (&beta)[MT],
                                                   int nk4, int mtc, unsigned long long mask) {
  if constexpr (MASKED) {
    if (mask != ~0ull) { mma_chunk_masked<MT, NT, KC>(acc, A, astride, B, beta, mtc, mask); return; }
  }
  if (mtc == MT) { mma_chunk<MT, NT, MT, KC, BETA>(acc, A, astride, B, beta, nk4); return; }
  if constexpr (MT > 1) { if (mtc == 1) { mma_chunk<MT, NT, 1, KC, BETA>(acc, A, astride, B, beta, nk4); return; } }
  if constexpr (MT > 2) { if (mtc == 2) { mma_chunk<MT, NT, 2, KC, BETA>(acc, A, astride, B, beta, nk4); return; } }
  if constexpr (MT > 3) { if (mtc == 3) { mma_chunk<MT, NT, 3, KC, BETA>(acc, A, astride, B, beta, nk4); return; } }
}

// Position in the CTA's chunk sequence (tile, block, column block, chunk); chunk == nchunks is the epilogue-constants chunk.
// Lives in shared memory: only warp 0 (the warp that re-arms the TMA ring) touches it.
struct Cursor {
  long long u;     // first unit of the tile being loaded
  int q, done;     // index into the per-tile TMA schedule
};

// barrier over the two warps (wn = 0, 1) that share a set of rows: everything a warp pair needs between the TMA ring and
// its Gram flush (parameters, coefficients, A rows, Gram copies) is produced by the pair itself, so pairs drift freely
// and one pair's tile prologue / epilogue overlaps the other pairs' DMMAs.
__device__ __forceinline__ void pair_barrier(int wm) { asm volatile("bar.sync %0, 64;" ::"r"(wm + 1) : "memory"); }

__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* addr, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}

#ifdef DLB_PROFILE
#define PROF_DECL long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long prof_c = clock64();
#define PROF(i) { const long long t_ = clock64(); prof_t[i] += t_ - prof_c; prof_c = t_; }
#define PROF_FLUSH if (lane == 0 && p.prof) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd((unsigned long long*)p.prof + i_, (unsigned long long)prof_t[i_]); }
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH
#endif

// MASKED: the launch has sub-tile masks (BlockDev::tmask); plans without any use the instance that carries no masked code at all (the masked
// path costs registers: C2 is 3 % slower with it compiled in, even when it never runs).
// QF: quadratic-form blocks (single accumulator per column; the epilogue multiplies every column by its feature and scatters weighted group sums
// into the Gram entries, see include/desilike_b200.h: dlb_block_desc::q_*).
template <int R, int MT, int NT, int WM, int CPS, int KC, bool RESIDENT, bool MASKED, bool QF = false>
__global__ void __launch_bounds__(WM * 64, CPS) contract_kernel(const __grid_constant__ ContractParams p) {
  constexpr int NW = 2 * WM, NTH = NW * 32, UNIT = WM * 8, BM = UNIT * MT, BN = 16 * NT;
  constexpr int STAGE_DOUBLES = BN * KC + (RESIDENT ? 0 : BM * KC);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = p.nstages;
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + (size_t)S * STAGE_DOUBLES);
  unsigned* cnt = reinterpret_cast<unsigned*>(full + S);              // [S] warps that released the stage (the last one re-arms it)
  uint64_t* abar = full + 2 * S;                                      // [WM] resident A rows copied from the feature matrix (per warp pair)
  Cursor* cur = reinterpret_cast<Cursor*>(abar + WM);
  double* A_s = reinterpret_cast<double*>(abar + WM + 6);             // [BM][ast]           (RESIDENT)
  double* xs0 = A_s + (RESIDENT ? (size_t)BM * p.ast : 0);            // [2][BM][dpad]       parameter tiles (double buffered)
  double* cs = xs0 + (size_t)2 * BM * p.dpad;                         // [ncoef][BM]         bias monomials
  double* red = cs + (size_t)(p.ncoef > 0 ? p.ncoef : 1) * BM;        // [2][BM][npart_max]  Gram partials (per warp column)
  GenDev* gens_s = reinterpret_cast<GenDev*>(red + (size_t)2 * BM * p.npart_max);   // [ngen]   (RESIDENT)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lc = lane & 3;
  const int ptid = wn * 32 + lane;        // thread index inside the warp pair
  const long long U = p.U;
  const long long u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  // Tile schedule: MT units per tile; with two CTAs per SM the second-resident CTA starts with a half tile, so that the two
  // CTAs of an SM run out of phase and one's tile prologue / epilogue overlaps the other's DMMAs.
  const bool half_first = CPS == 2 && MT > 1 && blockIdx.x >= (gridDim.x + 1) / 2;
  auto tile_units = [&](long long u) { return (int)min((long long)((half_first && u == u0) ? MT / 2 : MT), u1 - u); };

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); cnt[s] = 0u; }
    for (int i = 0; i < WM; ++i) mbar_init(&abar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cur->u = u0; cur->q = 0; cur->done = 0;
  }
  if constexpr (RESIDENT) {
    for (int i = tid; i < p.ngen * (int)(sizeof(GenDev) / 4); i += NTH) reinterpret_cast<int*>(gens_s)[i] = reinterpret_cast<const int*>(p.gens)[i];
  }
  if constexpr (QF && CPS == 2) {   // two CTAs per SM: start the second-resident one half a column block late, so that one's epilogue overlaps the other's DMMAs
    if (blockIdx.x >= (gridDim.x + 1) / 2) __nanosleep((unsigned)p.qf_offset_ns);
  }
  if constexpr (QF) {   // the epilogue reads feature columns up to the 8-column padding of a group (x exact zeros): every A entry must be finite
    for (int i = tid; i < BM * p.ast; i += NTH) A_s[i] = 0.;
    for (int i = tid; i < 2 * BM * p.dpad; i += NTH) xs0[i] = 0.;
    __syncthreads();
  }
  for (int i = tid; i < 2 * BM; i += NTH) xs0[(size_t)i * p.dpad + p.d] = 1.;   // constant-one column (absent monomial factors)
  // Parameter rows of this warp pair for a tile starting at unit u with mtc units -> buffer xb (cp.async; rows beyond N replicate
  // the last point, their results are never read)
  auto load_params = [&](long long u, int mtc, double* xb) {
    for (int i = ptid; i < mtc * 8 * p.d; i += 64) {
      const int rr = i / p.d, j = i - rr * p.d;
      const int r = (rr >> 3) * UNIT + wm * 8 + (rr & 7);
      cp_async8(xb + (size_t)r * p.dpad + j, p.params + min(u * UNIT + r, p.N - 1) * p.d + j);
    }
  };
  load_params(u0, tile_units(u0), xs0);
  __syncthreads();

  // ---- TMA ring: stage rs has been released by every warp (or is fresh); load the next transaction of the sequence into it.
  //      Executed by one whole warp (the last one to release the stage); the sequence position lives in shared memory and the
  //      per-tile schedule (source pointer + bytes of every k-chunk / epilogue-constants tile) is a plan-time table. ----
  auto refill = [&](int rs) {
    if (cur->done) return;
    const int q = cur->q;
    const long long cu = cur->u;
    double* dst = stage_base + (size_t)rs * STAGE_DOUBLES;
    const SchedEntry* e = p.sched + q;
    if constexpr (RESIDENT) {
      if (lane == 0) {
        const int4 ev = __ldg(reinterpret_cast<const int4*>(e));
        const double* src = reinterpret_cast<const double*>(((unsigned long long)(unsigned)ev.y << 32) | (unsigned)ev.x);
        mbar_arrive_expect_tx(&full[rs], (unsigned)ev.z);
        tma_bulk_g2s(dst, src, (unsigned)ev.z, &full[rs]);
      }
    } else {
      const int4 ev = __ldg(reinterpret_cast<const int4*>(e));
      const double* src = reinterpret_cast<const double*>(((unsigned long long)(unsigned)ev.y << 32) | (unsigned)ev.x);
      if (ev.w < 0) {   // epilogue constants
        if (lane == 0) { mbar_arrive_expect_tx(&full[rs], (unsigned)ev.z); tma_bulk_g2s(dst, src, (unsigned)ev.z, &full[rs]); }
      } else {
        const BlockDev* blk = p.blocks + (ev.w >> 16);
        const int4 ch = __ldg(reinterpret_cast<const int4*>(blk->chunks + (ev.w & 0xFFFF)));   // slot, coef, fcol, kc
        const int h = UNIT * tile_units(cu);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[rs], (uint32_t)(h * ch.w * 8) + (unsigned)ev.z);
          tma_bulk_g2s(dst, src, (unsigned)ev.z, &full[rs]);
        }
        __syncwarp();
        const double* Fsrc = p.F + cu * UNIT * p.fstride + ch.z;
        for (int r = lane; r < h; r += 32) tma_bulk_g2s(dst + BN * KC + r * KC, Fsrc + (size_t)r * p.fstride, ch.w * 8, &full[rs]);
      }
    }
    if (lane == 0) {   // advance
      if (q + 1 == p.nsched) {
        cur->q = 0;
        const long long nu = cu + tile_units(cu);
        cur->u = nu;
        if (nu >= u1) cur->done = 1;
      } else cur->q = q + 1;
      __threadfence_block();   // publish the position before this thread's next stage-release count
    }
    __syncwarp();
  };
  // this warp is done with stage rs: the last warp to say so re-arms it
  auto release = [&](int rs) {
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(&cnt[rs], 1u) == (unsigned)(NW - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      if (lane == 0) { cnt[rs] = 0u; __threadfence_block(); }
      __syncwarp();
      refill(rs);
    }
  };
  if (warp == 0)
    for (int i = 0; i < S; ++i) refill(i);

  const int npm = p.npart_max;
  const uint32_t full_u32 = smem_u32(full);
  int s = 0, xbuf = 0;
  uint32_t ph = 0, aph = 0;
  PROF_DECL
  for (long long u = u0; u < u1;) {
    const int mtc = tile_units(u);
    const long long row0 = u * UNIT;
    double* xs = xs0 + (size_t)xbuf * BM * p.dpad;
    cp_async_wait_all();
    pair_barrier(wm);   // this tile's parameter rows have landed; the pair has completely finished its previous tile
    PROF(5)
    if (u + mtc < u1) load_params(u + mtc, tile_units(u + mtc), xs0 + (size_t)(xbuf ^ 1) * BM * p.dpad);
    xbuf ^= 1;
    // ---- bias monomials c = scale * prod_f theta[idx_f]^pow_f (tracer combine, reference full_shape.py:545-550) ----
    for (int i = ptid; i < p.ncoef * mtc * 8; i += 64) {
      const int j = i / (mtc * 8), rr = i - j * (mtc * 8);
      const int r = (rr >> 3) * UNIT + wm * 8 + (rr & 7);
      const DevCoef* cf = p.coefs + j;
      double c = cf->scale;
      const int nf = cf->nf;
      for (int f = 0; f < nf; ++f) {
        const double x = xs[r * p.dpad + cf->idx[f]];
        const int e = cf->pow[f];
        c *= e >= 0 ? ipow(x, e) : 1. / ipow(x, -e);
      }
      cs[j * BM + r] = c;
    }
    for (int b = 0; b < p.nblocks; ++b) {
      const BlockDev* blk = p.blocks + b;
      const int nchunks = blk->nchunks, ncb = blk->ncb, npart = blk->npart, Rb = blk->nslots, nconst = blk->nconst;
      const Chunk* chunks = blk->chunks;
      PROF(5)
      if (b > 0) pair_barrier(wm);   // the pair's flush of the previous block is done and both warps finished reading their A rows
      PROF(5)
      if (RESIDENT && blk->fill == 1) {
        // ---- resident A rows = this block's generator segment of the feature matrix (MLP activations, ...): one TMA row copy
        //      per point, once per (tile, block) ----
        const int acols = blk->acols;
        if (ptid == 0) mbar_arrive_expect_tx(&abar[wm], (uint32_t)(mtc * 8 * acols * 8));
        pair_barrier(wm);
        if (ptid < mtc * 8) {
          const int r = (ptid >> 3) * UNIT + wm * 8 + (ptid & 7);
          tma_bulk_g2s(A_s + (size_t)r * p.ast, p.F + (row0 + r) * p.fstride + blk->fbase, acols * 8, &abar[wm]);
        }
        mbar_wait(&abar[wm], aph);
        aph ^= 1;
      } else if constexpr (RESIDENT) {
        // ---- Taylor monomials of this block's emulator, straight into shared memory ----
        //   m_t = prod_j (theta_j - center_j)^powers[t, j]  (layout: reference emulators/__init__.py:471-507).  paths[col] lists the
        //   factors of column col as bytes (variable index + 1, 0 terminates); top byte 0xFF marks a structural zero column.
        //   Every factor is read as xs[row][col_k] - c_k; absent factors point at the constant-one column with c_k = 0.
        const GenDev* g = gens_s + blk->gen;
        const int maxdeg = g->maxdeg;
        if (blk->mono2) {
          // two-pass form: columns of degree <= 2 from their two factors, then every column of degree 3 - 4 as ONE product of two
          // columns of the same row.  A warp only reads rows it wrote itself (rows wn, wn + 2, ...): __syncwarp orders the passes.
          const unsigned long long* m2 = blk->mono2;
          const int n1 = blk->n1, n2 = blk->n2;
          for (int e0 = 0; e0 < n1; e0 += 32) {
            const int e = e0 + lane;
            if (e < n1) {
              const unsigned long long w = __ldg(m2 + e);
              const int ca = (int)(w & 0xFFull), cb = (int)((w >> 8) & 0xFFull), col = (int)((w >> 32) & 0xFFFFull);
              const double va = g->cval[(w >> 16) & 0xFFull], vb = g->cval[(w >> 24) & 0xFFull];
              const bool zero = (w >> 56) != 0ull;
              for (int mt = 0; mt < mtc; ++mt) {
                const double* xb = xs + (size_t)(mt * UNIT + wm * 8 + wn) * p.dpad;
                double* ab = A_s + (size_t)(mt * UNIT + wm * 8 + wn) * p.ast + col;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const double* xr = xb + (size_t)(2 * i) * p.dpad;
                  ab[(size_t)(2 * i) * p.ast] = zero ? 0. : (xr[ca] - va) * (xr[cb] - vb);
                }
              }
            }
          }
          __syncwarp();
          for (int e0 = 0; e0 < n2; e0 += 32) {
            const int e = e0 + lane;
            if (e < n2) {
              const unsigned long long w = __ldg(m2 + n1 + e);
              const int cl = (int)(w & 0xFFFFull), cr = (int)((w >> 16) & 0xFFFFull), col = (int)((w >> 32) & 0xFFFFull);
              for (int mt = 0; mt < mtc; ++mt) {
                double* ab = A_s + (size_t)(mt * UNIT + wm * 8 + wn) * p.ast;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  double* ar = ab + (size_t)(2 * i) * p.ast;
                  ar[col] = ar[cl] * ar[cr];
                }
              }
            }
          }
        } else
        for (int j0 = 0; j0 < p.ast; j0 += 32) {
          const int j = j0 + lane;
          if (j < p.ast) {
            const ulonglong2 fw = __ldg(reinterpret_cast<const ulonglong2*>(blk->paths) + j);
            int col[7];
            double ck[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              col[k] = (int)((fw.x >> (8 * k)) & 0xFFull);
              ck[k] = g->cval[(fw.y >> (8 * k)) & 0xFFull];
            }
            const bool zero = (fw.x >> 56) != 0ull;
            for (int mt = 0; mt < mtc; ++mt) {
              const double* xb = xs + (size_t)(mt * UNIT + wm * 8 + wn) * p.dpad;
              double* ab = A_s + (size_t)(mt * UNIT + wm * 8 + wn) * p.ast + j;
#pragma unroll
              for (int i = 0; i < 4; ++i) {   // rows wn, wn + 2, wn + 4, wn + 6 of this pair's 8-row strip
                const double* xr = xb + (size_t)(2 * i) * p.dpad;
                double m = xr[col[0]] - ck[0];
#pragma unroll
                for (int k = 1; k < 7; ++k)
                  if (k < maxdeg) m *= xr[col[k]] - ck[k];
                ab[(size_t)(2 * i) * p.ast] = zero ? 0. : m;
              }
            }
          }
        }
      }
      if (p.mode == 0) {   // zero this warp's rows of its Gram copy
        for (int mt = 0; mt < mtc; ++mt)
          for (int i = lane; i < 8 * npart; i += 32) red[((size_t)wn * BM + mt * UNIT + wm * 8 + i / npart) * npm + (i % npart)] = 0.;
      }
      PROF(2)
      pair_barrier(wm);
      PROF(5)

      const int* cb_first = blk->cb_first;
      for (int cb = 0; cb < ncb; ++cb) {
        const int cfirst = cb_first ? __ldg(cb_first + cb) : 0;
        int4 ch_next = __ldg(reinterpret_cast<const int4*>(chunks + (cfirst < nchunks ? cfirst : 0)));   // slot, coef, fcol, kc (prefetched one chunk ahead)
        [[maybe_unused]] const unsigned long long* tmask = (MASKED && blk->tmask) ? blk->tmask + ((size_t)cb * nchunks) * 2 + wn : nullptr;   // [ncb][nchunks][2]
        double acc[R][MT][NT][2];
#pragma unroll
        for (int sl = 0; sl < R; ++sl)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[sl][mt][nt][0] = acc[sl][mt][nt][1] = 0.;

        for (int c = cfirst; c < nchunks; ++c) {
          const int4 ch = ch_next;
          unsigned long long mask = ~0ull;
          if constexpr (MASKED) { if (tmask) mask = __ldg(tmask + 2 * c); }   // (L1-resident table; the load overlaps the stage wait below)   // (L1-resident table; the load overlaps the stage wait below)
          ch_next = __ldg(reinterpret_cast<const int4*>(chunks + (c + 1 < nchunks ? c + 1 : c)));
          double beta[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) beta[mt] = ch.y >= 0 ? cs[ch.y * BM + mt * UNIT + wm * 8 + lr] : 1.;
          PROF(3)
          mbar_wait_u32(full_u32 + 8u * s, ph);
          PROF((c == cfirst ? (cb == 0 ? 1 : 7) : 4))
          const double* Bp = stage_base + (size_t)s * STAGE_DOUBLES + (wn * 8 + lr) * KC + lc;
          const double* Ap;
          int astride;
          if constexpr (RESIDENT) { Ap = A_s + (wm * 8 + lr) * p.ast + ch.z + lc; astride = UNIT * p.ast; }
          else { Ap = stage_base + (size_t)s * STAGE_DOUBLES + BN * KC + (wm * 8 + lr) * KC + lc; astride = UNIT * KC; }
          const int nk4 = ch.w >> 2;
#pragma unroll
          for (int sl = 0; sl < R; ++sl)
            if (ch.x == sl) mma_chunk_dispatch<MT, NT, KC, MASKED, !QF>(acc[sl], Ap, astride, Bp, beta, nk4, mtc, mask);
          PROF(3)
          release(s);
          PROF(5)
          if (++s == S) { s = 0; ph ^= 1; }
        }

        // ===== epilogue: r0 of this column block arrives as one more transaction of the ring (quadratic-form blocks have no such constants: their
        //       schedule carries the k-chunks only, so the next column block's chunk is already in flight) =====
        if constexpr (!QF) mbar_wait_u32(full_u32 + 8u * s, ph);
        PROF(0)
        const int colw = wn * 8 + 2 * lc;  // + nt * 16 + e   (interleaved n-tiles: warp wn owns the 8-column tiles 2 nt + wn)
        const double* epi = stage_base + (size_t)s * STAGE_DOUBLES;
        const double* r0 = epi;
        if (p.mode == 1) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < mtc) {
              const long long row = row0 + mt * UNIT + wm * 8 + lr;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int col = cb * BN + colw + nt * 16 + e;
                  if (row < p.out_rows && col < blk->ncols) {
                    double v = acc[0][mt][nt][e] + r0[colw + nt * 16 + e];
                    if (p.act == 1) v = silu(v);                   // silu  (reference emulators/conversion.py:27)
                    else if (p.act == 2) v = fmax(v, 0.);          // relu  (:29)
                    else if (p.act == 3) v = tanh(v);              // tanh  (:31)
                    double dscale = 1.;                            // u'(t) of the transformed columns
                    if (blk->tdata) {   // observable transform (reference observables/galaxy_clustering/power_spectrum.py:402-404)
                      const double dcol = __ldg(blk->tdata + col);
                      if (dcol != 0.) {
                        const double c = cbrt(v / dcol);
                        v = (3. * c - 2.) * dcol;
                        dscale = 1. / (c * c);
                      }
                    }
                    double* o = p.out + row * p.ostride + blk->ocol0 + col;
                    *o = v;
                    if constexpr (R > 1) {
                      // solved-parameter templates of a transformed block: Jacobian rows of u(t(x)) at the defaults, u'(t) dt/dx_a (the reference
                      // differentiates flatdiff THROUGH the transform, likelihoods/base.py:152-158)
                      if (blk->oslot > 0) {
#pragma unroll
                        for (int a = 1; a < R; ++a)
                          if (a < Rb) o[a * blk->oslot] = acc[a][mt][nt][e] * dscale;
                      }
                    }
                  }
                }
            }
          }
        } else if constexpr (QF) {
          // ===== quadratic-form epilogue: this warp's n-tiles of 8 columns each belong to one group.  Per tile one 16-byte record
          //       { group | valid columns << 16 | kind << 24, first feature column, Gram entry, (coef a + 1) | (coef b + 1) << 16 }.
          //       Every quadratic tile is reduced and scattered on its own -- short straight-line code (the version that merged the tiles of a
          //       group before reducing was if-converted into ~700 instructions per column block) =====
          const int4* qt = reinterpret_cast<const int4*>(blk->qtile) + (cb * (BN / 8) + wn);   // n-tile 2 nt + wn of the column block
          int4 meta[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) meta[nt] = __ldg(qt + 2 * nt);
          const double* frow[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) frow[mt] = A_s + (size_t)(mt * UNIT + wm * 8 + lr) * p.ast + 2 * lc;
          int any_direct = 0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            // columns beyond the group's K features hold exact zeros in the accumulators (zero matrix columns) and finite values in A_s (zeroed at
            // kernel start; the plan sizes the row stride so that whole 8-column tiles stay inside the row): no guard needed.  Padding tiles (group -1)
            // and direct tiles carry weight 0 here.
            const int kind = meta[nt].x >> 24, ca = (meta[nt].w & 0xFFFF) - 1, cbq = (meta[nt].w >> 16) - 1;
            const bool quad = kind == 0 && (meta[nt].x & 0xFFFF) != 0xFFFF;
            any_direct |= kind;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const double2 f = *reinterpret_cast<const double2*>(frow[mt] + meta[nt].y);
              double sq = fma(acc[0][mt][nt][0], f.x, acc[0][mt][nt][1] * f.y);
              sq += __shfl_xor_sync(0xffffffffu, sq, 1);
              sq += __shfl_xor_sync(0xffffffffu, sq, 2);
              if (lc == 0 && mt < mtc && quad) {
                const int row = mt * UNIT + wm * 8 + lr;
                const double w = (ca >= 0 ? cs[ca * BM + row] : 1.) * (cbq >= 0 ? cs[cbq * BM + row] : 1.);
                red[((size_t)wn * BM + row) * npm + meta[nt].z] += sq * w;
              }
            }
          }
          if (any_direct) {   // (warp-uniform; one column block in ~14 holds the direct groups) every column is its own Gram entry
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              if ((meta[nt].x >> 24) == 0) continue;
              const int gid = meta[nt].x & 0xFFFF;
              const int* de = blk->qdent + gid * 24 + 6 * lc;   // per column (Gram entry, coefficient a, coefficient b)
              int ent[2][3];
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int q = 0; q < 3; ++q) ent[e][q] = __ldg(de + 3 * e + q);
              // (the columns of a direct tile feed DISTINCT Gram entries -- lowering.qform_tables packs them that way: no conflicts between the lanes)
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                if (mt < mtc) {
                  const int row = mt * UNIT + wm * 8 + lr;
                  double* redrow = red + ((size_t)wn * BM + row) * npm;
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (ent[e][0] >= 0) {
                      const double w = (ent[e][1] >= 0 ? cs[ent[e][1] * BM + row] : 1.) * (ent[e][2] >= 0 ? cs[ent[e][2] * BM + row] : 1.);
                      redrow[ent[e][0]] += acc[0][mt][nt][e] * w;
                    }
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double2 rr = *reinterpret_cast<const double2*>(r0 + colw + nt * 16);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { acc[0][mt][nt][0] -= rr.x; acc[0][mt][nt][1] -= rr.y; }
          }
          // auxiliary columns ncols .. ncols + nconst - 1 of the block hold slot_a . g_i (dot products with the constant whitened
          // Jacobian rows, produced by the contraction itself): move them to their Gram entries and clear them so that the
          // pairwise sums below only see data-vector columns
          const int nS = (Rb * (Rb + 1)) / 2;
          for (int i = 0; i < nconst; ++i) {
            const int within = blk->ncols + i - cb * BN;
            if (within < 0 || within >= BN) continue;
            const int jt = within >> 3, wni = jt & 1, nti = jt >> 1, lci = (within & 7) >> 1, ei = within & 1;
            const bool mine = wn == wni && lc == lci;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (nt == nti && e == ei) {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                      if (mine) {   // only the lane that owns this column: the same register index is a data column elsewhere
                        if (mt < mtc && a < Rb) red[((size_t)wn * BM + mt * UNIT + wm * 8 + lr) * npm + nS + i * Rb + a] += acc[a][mt][nt][e];
                        acc[a][mt][nt][e] = 0.;
                      }
                    }
                }
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < mtc) {
              double* redrow = red + ((size_t)wn * BM + mt * UNIT + wm * 8 + lr) * npm;
#pragma unroll
              for (int a = 0; a < R; ++a) {
#pragma unroll
                for (int bb = a; bb < R; ++bb) {
                  if (bb < Rb) {
                    double sum = 0.;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) sum += acc[a][mt][nt][0] * acc[bb][mt][nt][0] + acc[a][mt][nt][1] * acc[bb][mt][nt][1];
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                    if (lc == 0) redrow[a * Rb - (a * (a - 1)) / 2 + (bb - a)] += sum;
                  }
                }
              }
            }
          }
        }
        PROF(6)
        if constexpr (!QF) {
          release(s);
          PROF(5)
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }  // cb

      if (p.mode == 0) {
        pair_barrier(wm);   // both warps' contributions are in their Gram copies
        PROF(5)
        double* out = p.partial + blk->poff * p.npad + row0;
        for (int i = ptid; i < mtc * 8 * npart; i += 64) {
          const int e = i / (mtc * 8), rr = i - e * (mtc * 8);
          const int r = (rr >> 3) * UNIT + wm * 8 + (rr & 7);
          double v = red[(size_t)r * npm + e] + red[((size_t)BM + r) * npm + e];
          if constexpr (QF) {
            v += __ldg(blk->qconst + e);
            for (int t = __ldg(blk->qw_off + e); t < __ldg(blk->qw_off + e + 1); ++t) {   // constant-feature parts: theta-dependent through their weight only
              const int ca = __ldg(blk->qw_coef + 2 * t), cbq = __ldg(blk->qw_coef + 2 * t + 1);
              v += __ldg(blk->qw_val + t) * (ca >= 0 ? cs[ca * BM + r] : 1.) * (cbq >= 0 ? cs[cbq * BM + r] : 1.);
            }
          }
          out[(long long)e * p.npad + r] = v;
        }
      }
      PROF(5)
    }  // blocks
    u += mtc;
  }    // tiles
  PROF_FLUSH
}

template <int R, int MT, int NT, int WM, int CPS, int KC, bool RESIDENT>
int launch_variant(const ContractParams& p, int grid, size_t smem, cudaStream_t stream) {
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    DLB_CUDA(cudaFuncSetAttribute(contract_kernel<R, MT, NT, WM, CPS, KC, RESIDENT, kMasked>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)contract_smem_limit(CPS)));
    configured = true;
  }
  contract_kernel<R, MT, NT, WM, CPS, KC, RESIDENT, kMasked><<<grid, WM * 64, smem, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

#if !DLB_CONTRACT_MASKED
template <int NT, int WM, int CPS, int KC>
int launch_qf(const ContractParams& p, int grid, size_t smem, cudaStream_t stream) {
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    DLB_CUDA(cudaFuncSetAttribute(contract_kernel<1, 1, NT, WM, CPS, KC, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)contract_smem_limit(CPS)));
    configured = true;
  }
  contract_kernel<1, 1, NT, WM, CPS, KC, true, false, true><<<grid, WM * 64, smem, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}
#endif

template <int R, int MT, int NT, int WM, int CPS, int KC>
int launch_mode(const ContractParams& p, int grid, size_t smem, cudaStream_t stream) {
  return p.resident ? launch_variant<R, MT, NT, WM, CPS, KC, true>(p, grid, smem, stream) : launch_variant<R, MT, NT, WM, CPS, KC, false>(p, grid, smem, stream);
}

}  // namespace

#if !DLB_CONTRACT_MASKED
size_t contract_smem_limit(int ctas_per_sm) {   // dynamic shared memory available to one CTA
  return ctas_per_sm == 2 ? (size_t)(233472 / 2 - 1024 - 512) : (size_t)227 * 1024;
}

size_t contract_smem_bytes(const Variant& v, int nstages, bool resident, int ast, int dpad, int ncoef, int npart_max, int ngen) {
  const int BM = variant_bm(v), BN = variant_bn(v), KC = v.KC;
  const size_t stage = (size_t)(BN * KC + (resident ? 0 : BM * KC)) * 8;
  size_t bytes = nstages * stage + (size_t)(2 * nstages + v.WM + 6) * 8;
  if (resident) bytes += (size_t)BM * ast * 8 + (size_t)ngen * sizeof(GenDev);
  bytes += (size_t)2 * BM * dpad * 8 + (size_t)std::max(ncoef, 1) * BM * 8 + (size_t)2 * BM * std::max(npart_max, 1) * 8;
  return bytes;
}

int launch_contract_masked(const Variant& v, const ContractParams& p, int nsm, cudaStream_t stream);   // contract_masked.cu
#endif

#if DLB_CONTRACT_MASKED
int launch_contract_masked(const Variant& v, const ContractParams& p, int nsm, cudaStream_t stream) {
#else
int launch_contract(const Variant& v, const ContractParams& p, int nsm, cudaStream_t stream) {
  if (p.masked && !p.qform) return launch_contract_masked(v, p, nsm, stream);
#endif
  const size_t smem = contract_smem_bytes(v, p.nstages, p.resident != 0, p.ast, p.dpad, p.ncoef, p.npart_max, p.ngen);
  if (smem > contract_smem_limit(v.CPS)) { set_error("contraction kernel: shared-memory budget exceeded"); return DLB_ERR_UNSUPPORTED; }
  const int grid = (int)std::min<long long>((long long)nsm * v.CPS, std::max<long long>(p.U, 1));
#if !DLB_CONTRACT_MASKED
  if (p.qform) {   // quadratic-form blocks: one accumulator per column, resident A; 16 warps x 64-row tiles, or 2 CTAs / SM x 8 warps x 32-row tiles
    if (!(v.R == 1 && v.MT == 1 && p.resident)) { set_error("quadratic-form plan: unexpected kernel variant"); return DLB_ERR_UNSUPPORTED; }
    if (v.WM == 8 && v.CPS == 1 && v.NT == 7 && v.KC == 36) return launch_qf<7, 8, 1, 36>(p, grid, smem, stream);
    if (v.WM == 8 && v.CPS == 1 && v.NT == 6 && v.KC == 76) return launch_qf<6, 8, 1, 76>(p, grid, smem, stream);
    if (v.WM == 8 && v.CPS == 1 && v.NT == 7 && v.KC == 68) return launch_qf<7, 8, 1, 68>(p, grid, smem, stream);
    if (v.WM == 4 && v.CPS == 2 && v.NT == 7 && v.KC == 36) return launch_qf<7, 4, 2, 36>(p, grid, smem, stream);
    set_error("quadratic-form plan: unexpected kernel variant");
    return DLB_ERR_UNSUPPORTED;
  }
#endif
  if (v.R == 1 && v.MT == 2 && v.NT == 7 && v.WM == 4 && v.CPS == 2 && v.KC == 28) return launch_mode<1, 2, 7, 4, 2, 28>(p, grid, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 7 && v.WM == 4 && v.CPS == 2 && v.KC == 36) return launch_mode<1, 2, 7, 4, 2, 36>(p, grid, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 5 && v.WM == 4 && v.CPS == 2 && v.KC == 28) return launch_mode<1, 2, 5, 4, 2, 28>(p, grid, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 5 && v.WM == 4 && v.CPS == 2 && v.KC == 36) return launch_mode<1, 2, 5, 4, 2, 36>(p, grid, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 7 && v.WM == 6 && v.CPS == 1 && v.KC == 76) return launch_mode<1, 2, 7, 6, 1, 76>(p, grid, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 7 && v.WM == 8 && v.CPS == 1 && v.KC == 36) return launch_mode<1, 2, 7, 8, 1, 36>(p, grid, smem, stream);
  if (v.R == 4 && v.MT == 1 && v.NT == 7 && v.WM == 6 && v.CPS == 1 && v.KC == 36) return launch_mode<4, 1, 7, 6, 1, 36>(p, grid, smem, stream);
  if (v.R == 4 && v.MT == 1 && v.NT == 7 && v.WM == 6 && v.CPS == 1 && v.KC == 76) return launch_mode<4, 1, 7, 6, 1, 76>(p, grid, smem, stream);
  if (v.R == 4 && v.MT == 1 && v.NT == 4 && v.WM == 8 && v.CPS == 1 && v.KC == 76) return launch_mode<4, 1, 4, 8, 1, 76>(p, grid, smem, stream);
  if (v.R == 8 && v.MT == 1 && v.NT == 4 && v.WM == 6 && v.CPS == 1 && v.KC == 36) return launch_mode<8, 1, 4, 6, 1, 36>(p, grid, smem, stream);
  if (v.R == 8 && v.MT == 1 && v.NT == 4 && v.WM == 6 && v.CPS == 1 && v.KC == 76) return launch_mode<8, 1, 4, 6, 1, 76>(p, grid, smem, stream);
  set_error("no contraction kernel variant for requested tile shape");
  return DLB_ERR_UNSUPPORTED;
}

}  // namespace dlb
