// Fused MLP emulator forward pass (sm_100a): x-operation -> every hidden layer v <- act(v W_l + b_l) with the activations kept
// in shared memory, one kernel for all MLP emulators of the likelihood.
//
// What it replaces: cosmoprimo's MLPEmulatorEngine.predict as driven by the reference's EmulatedCalculator
//   (reference desilike/emulators/__init__.py:408-409; operation vocabulary desilike/emulators/conversion.py:18-33,76-84:
//    x-operation '(v - lo) / (hi - lo)', dense 'v @ kernel + bias', activations silu / relu / tanh).
// The LAST linear layer and the inverse y-operations are affine in the last hidden activations and are folded into the
// contraction's term matrices at plan time (desilike_b200/mlp.py), so this kernel's output is the feature segment
// (1, 0, h_last) consumed by contract_kernel.
//
// Hardware mapping: persistent CTAs of 16 warps = two independent groups of 8 (each walks its own 64-row tiles between its own activation buffers,
// synchronised by a named barrier, started half a layer apart: one group's activation epilogue -- scalar FP64 -- overlaps the other's DMMAs; 8 warps
// alone are 2 per scheduler, which cannot saturate the DMMA issue rate); for each emulator the weights of all hidden layers are staged once per CTA in
// shared memory in the transposed, padded layout of a DMMA B operand (row stride = 4 mod 8: conflict-free 64-bit fragment loads);
// each 64-row tile then runs layer after layer on the FP64 tensor pipe (DMMA.8x8x4, warp tile 16 x 32) between two
// shared-memory activation buffers -- hidden activations never touch HBM (the per-layer launches of the generic contraction
// kernel wrote and re-read them: 1.0 ms of C4's 6.3 ms).  Used when every hidden width is <= 64 and the weights fit in shared
// memory; wider networks keep the per-layer launches.
#include "common.h"

namespace dlb {

namespace {

constexpr int MLP_MAXGROUPS = 2, MLP_GTHREADS = 256, MLP_THREADS = MLP_MAXGROUPS * MLP_GTHREADS;   // up to two groups of 8 warps (one when shared memory is short)
constexpr int MLP_BM = 64, MLP_H = 64, MLP_AS = 68;   // rows per tile, max width, activation row stride (= 4 mod 8)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double activate(double v, int act) {
  if (act == 1) return v * fast_rcp(1. + fast_exp(fmin(fmax(-v, -708.), 708.)));   // silu (reference emulators/conversion.py:27)
  if (act == 2) return fmax(v, 0.);                                                // relu (:29)
  if (act == 3) return tanh(v);                                                    // tanh (:31)
  return v;
}

struct MlpFusedParams {
  const DevGen* gens; const int* mlp_index; int nmlp;   // generators table + indices of the MLP ones
  const double* params; int d; long long N; long long npad;
  double* F; int fstride;
  int groups;   // warp groups per CTA (1 or 2)
};

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_fused_kernel(const MlpFusedParams p) {
  extern __shared__ double sm[];
  const int grp = threadIdx.x / MLP_GTHREADS;   // warp group: its own tiles and activation buffers
  double* H0 = sm + (size_t)grp * 2 * MLP_BM * MLP_AS;   // [64][68] activations (ping)
  double* H1 = H0 + MLP_BM * MLP_AS;                      // [64][68] activations (pong)
  double* Wt = sm + (size_t)p.groups * 2 * MLP_BM * MLP_AS;   // per layer: [64][Hin_pad + 4] transposed weights, then bias [64] (shared by the groups)
  const int tid = threadIdx.x % MLP_GTHREADS, warp = tid >> 5, lane = tid & 31;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(MLP_GTHREADS) : "memory"); };
  const int wm = warp >> 1, wn = warp & 1, lr = lane >> 2, lc = lane & 3;
  const long long ntiles = (p.npad + MLP_BM - 1) / MLP_BM;
  for (int ig = 0; ig < p.nmlp; ++ig) {
    const DevGen& g = p.gens[p.mlp_index[ig]];
    const int nh = g.idata[0], act = g.idata[1], d = g.idata[2];
    __syncthreads();   // previous emulator's tiles done with Wt / H
    // ---- stage the weights: Wt_l[n][k] = W_l[k][n], zero padded to 64 outputs and Hin rounded up to 4 ----
    {
      const double* w = g.ddata + 2 * d;
      double* dst = Wt;
      int hin = d;
      for (int l = 0; l < nh; ++l) {
        const int hout = g.idata[3 + l], kpad = (hin + 3) & ~3, ks = kpad + 4;
        for (int i = threadIdx.x; i < MLP_H * ks; i += blockDim.x) {
          const int n = i / ks, k = i - n * ks;
          dst[i] = (n < hout && k < hin) ? w[(size_t)k * hout + n] : 0.;
        }
        const double* bias = w + (size_t)hin * hout;
        for (int n = threadIdx.x; n < MLP_H; n += blockDim.x) dst[MLP_H * ks + n] = n < hout ? bias[n] : 0.;
        dst += MLP_H * ks + MLP_H;
        w = bias + hout;
        hin = hout;
      }
    }
    __syncthreads();
    const int hlast = g.idata[2 + nh];
    if (grp == 1) __nanosleep(2000);   // half a layer out of phase with group 0
    for (long long tile = (long long)blockIdx.x * p.groups + grp; tile < ntiles; tile += (long long)gridDim.x * p.groups) {
      const long long row0 = tile * MLP_BM;
      // ---- x-operation v * a + b (reference conversion.py:76-77) into H0, zero padded to a multiple of 4 ----
      const int dpad = (d + 3) & ~3;
      for (int i = tid; i < MLP_BM * dpad; i += MLP_GTHREADS) {
        const int r = i / dpad, j = i - r * dpad;
        double v = 0.;
        if (j < d) {
          const long long row = min(row0 + r, p.N - 1);
          const double x = g.pidx[j] >= 0 ? p.params[row * p.d + g.pidx[j]] : g.pdef[j];
          v = x * g.ddata[j] + g.ddata[d + j];
        }
        H0[r * MLP_AS + j] = v;
      }
      group_sync();
      double* Hin = H0;
      double* Hout = H1;
      const double* wl = Wt;
      int hin = d;
      for (int l = 0; l < nh; ++l) {
        const int kpad = (hin + 3) & ~3, ks = kpad + 4;
        const double* bias = wl + MLP_H * ks;
        double acc[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.;
        const double* A = Hin + (wm * 16 + lr) * MLP_AS + lc;
        const double* B = wl + (wn * 32 + lr) * ks + lc;
#pragma unroll 4
        for (int k4 = 0; k4 < kpad / 4; ++k4) {
          double a[2], b[4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[mt] = A[mt * 8 * MLP_AS + k4 * 4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) b[nt] = B[nt * 8 * ks + k4 * 4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        const bool last = l == nh - 1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r = wm * 16 + mt * 8 + lr;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int col = wn * 32 + nt * 8 + 2 * lc;
            const double v0 = activate(acc[mt][nt][0] + bias[col], act), v1 = activate(acc[mt][nt][1] + bias[col + 1], act);
            if (last) {   // feature segment (1, 0, h_last): written straight to the feature matrix
              const long long row = row0 + r;
              if (row < p.npad) {
                double* out = p.F + row * p.fstride + g.foff;
                if (col < hlast) out[2 + col] = v0;
                if (col + 1 < hlast) out[2 + col + 1] = v1;
                if (col == 0) { out[0] = row < p.N ? 1. : 0.; out[1] = 0.; }
              }
            } else {
              *reinterpret_cast<double2*>(Hout + r * MLP_AS + col) = make_double2(v0, v1);
            }
          }
        }
        group_sync();
        double* t = Hin; Hin = Hout; Hout = t;
        wl += MLP_H * ks + MLP_H;
        hin = g.idata[3 + l];
      }
    }
  }
}

}  // namespace

// shared memory of the fused kernel for one generator with `groups` warp groups (host mirror of the staging loop); 0 if the network does not qualify
size_t mlp_fused_smem(const int* idata_host, int groups) {
  const int nh = idata_host[0], d = idata_host[2];
  if (nh < 1 || d > MLP_H) return 0;
  size_t doubles = (size_t)groups * 2 * MLP_BM * MLP_AS;
  int hin = d;
  for (int l = 0; l < nh; ++l) {
    const int hout = idata_host[3 + l];
    if (hout > MLP_H) return 0;
    doubles += (size_t)MLP_H * (((hin + 3) & ~3) + 4) + MLP_H;
    hin = hout;
  }
  return doubles * sizeof(double);
}

int launch_mlp_fused(const DevGen* d_gens, const int* d_mlp_index, int nmlp, size_t smem, int groups, const double* params, int d, long long N, long long npad,
                     double* F, int fstride, int nsm, cudaStream_t stream) {
  static size_t configured_dev[64] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    DLB_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  MlpFusedParams p;
  p.gens = d_gens; p.mlp_index = d_mlp_index; p.nmlp = nmlp; p.params = params; p.d = d; p.N = N; p.npad = npad; p.F = F; p.fstride = fstride;
  p.groups = groups;
  const long long ntiles = (npad + MLP_BM - 1) / MLP_BM;
  mlp_fused_kernel<<<(unsigned)std::min<long long>(nsm, (ntiles + groups - 1) / groups), groups * MLP_GTHREADS, smem, stream>>>(p);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb
