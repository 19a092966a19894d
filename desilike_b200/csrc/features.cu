// Feature generators: theta[N, d] -> per-point feature rows F[Npad, fstride] (one padded segment per generator).  The Taylor kernel is
// only used by the debug entry dlb_features_batch and for orders > 7 (otherwise the monomials are evaluated inside contract_kernel).  Rows >= N and the padding columns are written as zeros so the
// contraction kernel can consume full tiles without masking.
#include "common.h"

namespace dlb {

namespace {

__device__ __forceinline__ double ipow(double x, int n) {
  double r = 1.;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Taylor monomials  m_t = prod_j (theta_j - center_j)^powers[t, j]
//   layout fixed by the reference's fit (reference desilike/emulators/__init__.py:471-507: center, powers[T, d];
//   1/alpha! is already folded into the stored derivatives); predict restated in oracle/core.py:taylor_monomials.
//   idata = powers [T, nparams];  ddata = center [nparams];  nfeat = 2 + T: features = (1, 0, m_0 .. m_{T-1}).
//   One thread per (point, term): coalesced stores.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void taylor_features_kernel(const DevGen g, const double* __restrict__ params, int d, long long N, long long npad,
                                       double* __restrict__ F, int fstride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p = idx / g.fpad;
  const int t = (int)(idx - p * g.fpad);
  if (p >= npad) return;
  double m = 0.;
  if (p < N && t == 0) m = 1.;  // every segment starts with the constant pair (1, 0)
  if (p < N && t >= 2 && t < g.nfeat) {
    m = 1.;
    const int* pw = g.idata + (size_t)(t - 2) * g.nparams;
    for (int j = 0; j < g.nparams; ++j) {
      const int e = pw[j];
      if (e) {
        const double x = (g.pidx[j] >= 0 ? params[p * d + g.pidx[j]] : g.pdef[j]) - g.ddata[j];
        m *= ipow(x, e);
      }
    }
  }
  F[p * fstride + g.foff + t] = m;
}

// ---------------------------------------------------------------------------------------------------------------------
// MLP input stage: the x-operation '(v - lo) / (hi - lo)' (reference emulators/conversion.py:76-77) as v * a + b, written
// into the generator's xs sub-segment (g.aux[0]); also sets the constant pair (1, 0).  ddata = [a (d), b (d), ...].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mlp_input_kernel(const DevGen g, const double* __restrict__ params, int d, long long N, long long npad,
                                 double* __restrict__ F, int fstride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = g.nparams + 2;
  const long long p = idx / w;
  const int j = (int)(idx - p * w);
  if (p >= npad) return;
  double* row = F + p * fstride + g.foff;
  if (j >= g.nparams) { row[j - g.nparams] = (j == g.nparams && p < N) ? 1. : 0.; return; }
  double v = 0.;
  if (p < N) {
    const double x = g.pidx[j] >= 0 ? params[p * d + g.pidx[j]] : g.pdef[j];
    v = x * g.ddata[j] + g.ddata[g.nparams + j];
  }
  row[g.aux[0] + j] = v;
}

}  // namespace

int launch_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream) {
  const int threads = 256;
  if (g.kind == DLB_GEN_TAYLOR) {
    const long long total = npad * g.fpad;
    taylor_features_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(g, params, d, N, npad, F, fstride);
    g_launches++;
    DLB_CUDA(cudaGetLastError());
    return DLB_OK;
  }
  return launch_theory_features(g, params, d, N, npad, F, fstride, stream);
}

// Derived-expression parameters (reference parameter.py:795-808, applied per point at base.py:537): one thread per point copies the d_in input
// columns and appends the nder derived values, each a reverse-Polish program over the row (opcodes: include/desilike_b200.h).
__global__ void derived_kernel(const double* __restrict__ params, int d_in, int nder, const int* __restrict__ off, const int* __restrict__ code,
                               const double* __restrict__ consts, long long N, double* __restrict__ out) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= N) return;
  const int d = d_in + nder;
  double* row = out + pt * d;
  for (int j = 0; j < d_in; ++j) row[j] = params[pt * d_in + j];
  for (int i = 0; i < nder; ++i) {
    double st[16];
    int sp = 0;
    for (int o = off[i]; o < off[i + 1]; ++o) {
      const int op = code[2 * o], arg = code[2 * o + 1];
      if (op == 0) { st[sp++] = row[arg]; continue; }
      if (op == 1) { st[sp++] = consts[arg]; continue; }
      if (op <= 6) {
        const double b = st[--sp], a = st[sp - 1];
        st[sp - 1] = op == 2 ? a + b : op == 3 ? a - b : op == 4 ? a * b : op == 5 ? a / b : pow(a, b);
        continue;
      }
      const double a = st[sp - 1];
      st[sp - 1] = op == 7 ? -a : op == 8 ? exp(a) : op == 9 ? log(a) : op == 10 ? sqrt(a) : op == 11 ? log10(a) : op == 12 ? fabs(a) :
                   op == 13 ? sin(a) : op == 14 ? cos(a) : tanh(a);
    }
    row[d_in + i] = st[0];
  }
}

int launch_derived(const double* params, int d_in, int nder, const int* off, const int* code, const double* consts, long long N, double* out,
                   cudaStream_t stream) {
  const int threads = 128;
  derived_kernel<<<(unsigned)((N + threads - 1) / threads), threads, 0, stream>>>(params, d_in, nder, off, code, consts, N, out);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

int launch_mlp_input(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream) {
  const int threads = 256;
  const long long total = npad * (g.nparams + 2);
  mlp_input_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(g, params, d, N, npad, F, fstride);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb
