// Fused contraction kernel (sm_100a): for a tile of parameter points and a block of data-vector columns
//
//     slot[s][p, :] = sum_{terms t of slot s} coef_t[p] * F[p, fcol_t : fcol_t + K_t] . M_t[K_t, cols]
//
// followed, in registers, by the chi2 / normal-equation (Gram) reduction over the columns.  This one kernel replaces,
// per point, the reference's   emulator tensordot -> tracer bias combine -> window matrix apply -> flatdiff -> d^T P d
// and the J P d, J P J^T products of the analytic marginalisation
//   (reference emulators/__init__.py:471-507; full_shape.py:545-550,628-634; window.py:459-473; likelihoods/base.py:13-17,
//    :165-174, :658-664): all constant matrices (Taylor derivatives / last MLP layer, window W, Cholesky factor L of the
//    precision) are folded into M_t at plan time, so no theory vector ever touches HBM.
//
// Hardware mapping
//   * FP64 tensor pipe: DMMA.8x8x4 (mma.sync.m8n8k4.f64) -- tcgen05 has no FP64 kind; every FP64 mma shape lowers to
//     DMMA.8x8x4 on sm_100a (checked with cuobjdump).  8 consumer warps in a 4 (m) x 2 (n) grid, warp tile (MT*8) x (NT*8).
//   * operands staged in shared memory by the TMA unit: 1-D bulk copies (cp.async.bulk -> UBLKCP) issued by a dedicated
//     producer warp into an mbarrier ring (full/empty), 2-4 stages; the term matrices are pre-tiled at plan time into
//     the exact shared-memory image (n-major rows of KC = 36 doubles: 36 = 4 mod 8 makes both 64-bit fragment loads
//     bank-conflict free), so one bulk copy moves a whole B tile and one per point row moves the A tile.
//   * per-point coefficient (bias monomial) applied to the A fragment in registers; the Gram epilogue is warp-shuffle
//     reductions + shared-memory atomics, then one coalesced store of the partial sums.
#include "common.h"

namespace dlb {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int R, int MT, int NT>
__global__ void __launch_bounds__(NTHREADS, 1) contract_kernel(const __grid_constant__ ContractParams p) {
  constexpr int BM = 4 * MT * 8, BN = 2 * NT * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = p.nstages;
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  constexpr int STAGE_DOUBLES = (BM + BN) * KC;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + (size_t)S * STAGE_DOUBLES);
  uint64_t* empty = full + S;
  Chunk* schunks = reinterpret_cast<Chunk*>(empty + S);
  double* red = reinterpret_cast<double*>(schunks + ((p.nchunks + 1) & ~1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.y;
  const long long row0 = (long long)blockIdx.x * BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < p.nchunks; c += NTHREADS) schunks[c] = p.chunks[c];
  if (p.mode == 0)
    for (int i = threadIdx.x; i < BM * p.npart; i += NTHREADS) red[i] = 0.;
  __syncthreads();

  if (warp == NCONSUMER_WARPS) {
    // ===== TMA producer warp =====
    const double* tiles = p.tiles + (size_t)cb * p.nchunks * (BN * KC);
    for (int c = 0; c < p.nchunks; ++c) {
      const int s = c % S;
      const uint32_t ph = (c / S) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const Chunk ch = schunks[c];
      double* As = stage_base + (size_t)s * STAGE_DOUBLES;
      double* Bs = As + BM * KC;
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[s], (uint32_t)(BM * ch.kc * 8 + BN * KC * 8));
        tma_bulk_g2s(Bs, tiles + (size_t)c * (BN * KC), BN * KC * 8, &full[s]);
      }
      __syncwarp();
      const double* Fsrc = p.F + row0 * p.fstride + ch.fcol;
      for (int r = lane; r < BM; r += 32) tma_bulk_g2s(As + r * KC, Fsrc + (size_t)r * p.fstride, ch.kc * 8, &full[s]);
    }
    return;
  }

  // ===== consumer warps =====
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lc = lane & 3;
  double acc[R][MT][NT][2];
#pragma unroll
  for (int s = 0; s < R; ++s)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[s][mt][nt][0] = acc[s][mt][nt][1] = 0.;

  const double* Crow = p.C + (row0 + wm * (MT * 8) + lr) * p.cstride;
  double beta_next[MT];
  {
    const int coef = schunks[0].coef;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) beta_next[mt] = coef >= 0 ? __ldg(Crow + (size_t)mt * 8 * p.cstride + coef) : 1.;
  }
  for (int c = 0; c < p.nchunks; ++c) {
    const int s = c % S;
    const uint32_t ph = (c / S) & 1;
    const Chunk ch = schunks[c];
    double beta[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) beta[mt] = beta_next[mt];
    if (c + 1 < p.nchunks) {  // prefetch the next chunk's coefficient while this chunk computes
      const int coef = schunks[c + 1].coef;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) beta_next[mt] = coef >= 0 ? __ldg(Crow + (size_t)mt * 8 * p.cstride + coef) : 1.;
    }
    mbar_wait(&full[s], ph);
    const double* A = stage_base + (size_t)s * STAGE_DOUBLES + (wm * (MT * 8) + lr) * KC + lc;
    const double* B = stage_base + (size_t)s * STAGE_DOUBLES + BM * KC + (wn * (NT * 8) + lr) * KC + lc;
    const int nk4 = ch.kc >> 2;
#pragma unroll
    for (int slot = 0; slot < R; ++slot) {
      if (ch.slot == slot) {
#pragma unroll 3
        for (int k4 = 0; k4 < nk4; ++k4) {
          double a[MT], b[NT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) a[mt] = A[mt * 8 * KC + k4 * 4] * beta[mt];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) b[nt] = B[nt * 8 * KC + k4 * 4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dmma884(acc[slot][mt][nt][0], acc[slot][mt][nt][1], a[mt], b[nt]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue =====
  const int colw = wn * (NT * 8) + 2 * lc;  // + nt * 8 + e
  const double* r0 = p.r0 + (size_t)cb * BN;
  if (p.mode == 1) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const long long row = row0 + wm * (MT * 8) + mt * 8 + lr;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = cb * BN + colw + nt * 8 + e;
          if (row < p.out_rows && col < p.ncols) {
            double v = acc[0][mt][nt][e] + r0[colw + nt * 8 + e];
            if (p.act == 1) v = v / (1. + exp(-v));        // silu  (reference emulators/conversion.py:27)
            else if (p.act == 2) v = fmax(v, 0.);          // relu  (:29)
            else if (p.act == 3) v = tanh(v);              // tanh  (:31)
            p.out[row * p.ostride + p.ocol0 + col] = v;
          }
        }
    }
    return;
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const double2 rr = *reinterpret_cast<const double2*>(r0 + colw + nt * 8);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { acc[0][mt][nt][0] -= rr.x; acc[0][mt][nt][1] -= rr.y; }
  }
  const int Rb = p.nslots;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double* redrow = red + (wm * (MT * 8) + mt * 8 + lr) * p.npart;
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
      for (int b = a; b < R; ++b) {
        if (b < Rb) {
          double s = 0.;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) s += acc[a][mt][nt][0] * acc[b][mt][nt][0] + acc[a][mt][nt][1] * acc[b][mt][nt][1];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (lc == 0) atomicAdd(redrow + (a * Rb - (a * (a - 1)) / 2 + (b - a)), s);
        }
      }
    }
    const int nS = (Rb * (Rb + 1)) / 2;
    for (int i = 0; i < p.nconst; ++i) {
      const double* g = p.gconst + ((size_t)i * gridDim.y + cb) * BN + colw;
      double gv[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(g + nt * 8));
        gv[nt][0] = t.x; gv[nt][1] = t.y;
      }
#pragma unroll
      for (int a = 0; a < R; ++a) {
        if (a < Rb) {
          double s = 0.;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) s += acc[a][mt][nt][0] * gv[nt][0] + acc[a][mt][nt][1] * gv[nt][1];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (lc == 0) atomicAdd(redrow + nS + i * Rb + a, s);
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
  double* out = p.partial + (p.poff + (long long)cb * p.npart) * p.npad + row0;
  for (int i = threadIdx.x; i < BM * p.npart; i += NCONSUMER_WARPS * 32) {
    const int e = i / BM, r = i - e * BM;
    out[(long long)e * p.npad + r] = red[r * p.npart + e];
  }
}

template <int R, int MT, int NT>
int launch_variant(const ContractParams& p, int ntiles, int ncb, size_t smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    DLB_CUDA(cudaFuncSetAttribute(contract_kernel<R, MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid(ntiles, ncb);
  contract_kernel<R, MT, NT><<<grid, NTHREADS, smem, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace

size_t contract_smem_bytes(const Variant& v, int nstages, int nchunks, int npart) {
  const size_t stage = (size_t)(variant_bm(v) + variant_bn(v)) * KC * 8;
  return nstages * stage + 2 * nstages * 8 + (size_t)((nchunks + 1) & ~1) * sizeof(Chunk) + (size_t)variant_bm(v) * npart * 8;
}

int launch_contract(const Variant& v, const ContractParams& p, int ntiles, int ncb, cudaStream_t stream) {
  const size_t smem = contract_smem_bytes(v, p.nstages, p.nchunks, p.mode == 0 ? p.npart : 0);
  if (v.R == 1 && v.MT == 4 && v.NT == 7) return launch_variant<1, 4, 7>(p, ntiles, ncb, smem, stream);
  if (v.R == 1 && v.MT == 2 && v.NT == 7) return launch_variant<1, 2, 7>(p, ntiles, ncb, smem, stream);
  if (v.R == 4 && v.MT == 2 && v.NT == 4) return launch_variant<4, 2, 4>(p, ntiles, ncb, smem, stream);
  if (v.R == 8 && v.MT == 1 && v.NT == 4) return launch_variant<8, 1, 4>(p, ntiles, ncb, smem, stream);
  set_error("no contraction kernel variant for requested tile shape");
  return DLB_ERR_UNSUPPORTED;
}

}  // namespace dlb
